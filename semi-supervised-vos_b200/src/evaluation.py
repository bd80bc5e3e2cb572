"""`evaluation` command: mean J, F and J&F of a tree of result PNGs against the ground-truth tree (the reference's
src/evaluation.py:16-75 -- same command name, options, return value and pairing rules; offline CPU tooling, not on the hot path).

Pairing rules kept from the reference, because its published numbers depend on them:
  * files are paired by position in the two sorted recursive listings, not by name;
  * a result is resized (PIL's default filter for 'P' images: nearest) to its ground truth's size;
  * inside a pair the k-th smallest palette index of the ground truth is scored against the k-th smallest of the result,
    background included, and surplus indices on either side are dropped;
  * a pair's score is the plain mean over those index pairs; the totals are plain means over the pairs.
Pairs are scored in worker processes, handed out in chunks so that the pool's queue traffic stays small on large trees."""
from __future__ import annotations

import multiprocessing as mp
from pathlib import Path
from typing import Iterable, List, Sequence, Tuple

import click
import numpy as np
from loguru import logger
from PIL import Image
from tqdm import tqdm

from src.config import Config
from src.utils.metrics import evaluate_segmentation


def _indexed(path, size=None) -> np.ndarray:
    """Palette-index image as an array, brought to `size` (W, H) when given."""
    img = Image.open(path).convert('P')
    if size is not None:
        img = img.resize(size)
    return np.asarray(img)


def process_pair(gt, seg) -> np.ndarray:
    """(J, F) of one result PNG against its ground truth: mean over the rank-matched palette indices."""
    truth = _indexed(gt)
    result = _indexed(seg, size=truth.shape[::-1])
    per_object = np.empty((0, 2))
    for t_idx, r_idx in zip(np.unique(truth), np.unique(result)):
        per_object = np.vstack([per_object, evaluate_segmentation(truth == t_idx, result == r_idx)])
    return per_object.mean(axis=0)


def _score_star(pair: Tuple[Path, Path]) -> np.ndarray:
    return process_pair(*pair)


def _listing(root) -> List[Path]:
    return sorted(Path(root).rglob('*.png'))


def score_tree(pairs: Sequence[Tuple[Path, Path]], workers: int, quiet: bool) -> np.ndarray:
    """(n_pairs, 2) array of per-pair (J, F), in the order of `pairs`."""
    chunk = max(1, len(pairs) // (8 * max(workers, 1)))
    with mp.Pool(workers) as pool:
        scored: Iterable[np.ndarray] = pool.imap(_score_star, pairs, chunksize=chunk)
        rows = list(tqdm(scored, total=len(pairs), disable=quiet))
    return np.asarray(rows, dtype=np.float64).reshape(len(pairs), 2)


@click.command(name='evaluation')
@click.option('--ground_truth', '-g', type=click.Path(file_okay=False, dir_okay=True), required=True,
              help='Path to ground truth dataset folder.')
@click.option('--computed_results', '-c', type=click.Path(file_okay=False, dir_okay=True), required=True,
              help='Path to computed results.')
def evaluation_command(ground_truth, computed_results):
    evaluation_command_impl(ground_truth, computed_results)


def evaluation_command_impl(ground_truth, computed_results, disable=False):
    truth_files, result_files = _listing(ground_truth), _listing(computed_results)
    if len(truth_files) != len(result_files):
        raise AssertionError(f'{len(truth_files)} ground-truth PNGs against {len(result_files)} results')
    logger.info(f'Scoring {len(truth_files)} result masks against their ground truth.')
    table = score_tree(list(zip(truth_files, result_files)), Config.CPU_COUNT, quiet=disable)
    j_mean, f_mean = (float(v) for v in table.mean(axis=0))
    jf_mean = 0.5 * (j_mean + f_mean)
    logger.info(f'Evaluated: j_mean={j_mean}, f_mean={f_mean}, j&f_mean={jf_mean}.')
    return j_mean, f_mean, jf_mean
