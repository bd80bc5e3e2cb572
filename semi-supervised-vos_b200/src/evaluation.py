"""`evaluation` command: mean J, F and J&F of a directory of result PNGs against the ground-truth PNGs (reference
src/evaluation.py:16-75).  Pairs are matched by sorted path; inside a pair the k-th colour of the ground truth is scored
against the k-th colour of the result (background included), as the reference does."""
from multiprocessing import Pool
from pathlib import Path

import click
import numpy as np
from loguru import logger
from PIL import Image
from tqdm import tqdm

from src.config import Config
from src.utils.metrics import evaluate_segmentation


def process_pair(gt, seg):
    gt_img = Image.open(gt).convert('P')
    seg_img = Image.open(seg).convert('P').resize(gt_img.size)
    gt_img, seg_img = np.asarray(gt_img), np.asarray(seg_img)
    scores = [evaluate_segmentation(gt_img == g, seg_img == s) for g, s in zip(np.unique(gt_img), np.unique(seg_img))]
    return np.array(scores).mean(axis=0)


@click.command(name='evaluation')
@click.option('--ground_truth', '-g', type=click.Path(file_okay=False, dir_okay=True), required=True,
              help='Path to ground truth dataset folder.')
@click.option('--computed_results', '-c', type=click.Path(file_okay=False, dir_okay=True), required=True,
              help='Path to computed results.')
def evaluation_command(ground_truth, computed_results):
    evaluation_command_impl(ground_truth, computed_results)


def evaluation_command_impl(ground_truth, computed_results, disable=False):
    ground_truth = sorted(Path(ground_truth).glob('**/*.png'))
    computed = sorted(Path(computed_results).glob('**/*.png'))
    assert len(ground_truth) == len(computed)
    logger.info(f'Staring evaluation on {len(ground_truth)} pairs.')
    pbar = tqdm(total=len(ground_truth), disable=disable)
    with Pool(Config.CPU_COUNT) as pool:
        jobs = [pool.apply_async(process_pair, args=(gt, seg), callback=lambda _: pbar.update(1))
                for gt, seg in zip(ground_truth, computed)]
        scores = np.array([job.get() for job in jobs])
    pbar.close()
    j_mean, f_mean = scores[:, 0].mean(), scores[:, 1].mean()
    jf_mean = np.array([j_mean, f_mean]).mean()
    logger.info(f'Evaluated: j_mean={j_mean}, f_mean={f_mean}, j&f_mean={jf_mean}.')
    return j_mean, f_mean, jf_mean
