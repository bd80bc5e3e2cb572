"""J / F measures and the `evaluation` command against values produced by the reference's own
src/utils/metrics.py and src/evaluation.py::process_pair (oracle/make_golden_eval.py; the F-measure's two scikit-image
primitives came from the scipy shim documented there)."""
import json

import numpy as np
from PIL import Image

from oracle import reference_harness as RH
from oracle.fixtures import mask_pair
from tests._golden import GOLDEN

G = json.loads((GOLDEN / 'eval_jf.json').read_text())


def _save(path, arr):
    img = Image.fromarray(arr, mode='P')
    img.putpalette(RH.default_palette())
    img.save(path)


def test_j_and_f_match_reference(tmp_path):
    from src.evaluation import process_pair
    from src.utils.metrics import evaluate_segmentation
    for case in G['pairs']:
        seed = case['seed']
        gt, seg = mask_pair(seed, n_obj=1 + seed % 3, jitter=1 + seed % 4)
        got = [[float(v) for v in evaluate_segmentation(gt == k, seg == k)] for k in range(int(gt.max()) + 1)]
        assert np.allclose(got, case['per_object'], rtol=0, atol=1e-12), seed
        _save(tmp_path / 'gt.png', gt)
        _save(tmp_path / 'seg.png', seg)
        assert np.allclose(process_pair(tmp_path / 'gt.png', tmp_path / 'seg.png'), case['process_pair'], rtol=0, atol=1e-12)


def test_corner_cases_and_stacks():
    from src.utils.metrics import eval_f, eval_j, evaluate_segmentation
    z, o = np.zeros((40, 60), bool), np.zeros((40, 60), bool)
    o[10:20, 10:30] = True
    for name, (a, b) in {'both_empty': (z, z), 'seg_empty': (o, z), 'gt_empty': (z, o)}.items():
        assert [float(v) for v in evaluate_segmentation(a, b)] == G['empty'][name], name
    stack = [mask_pair(100 + t) for t in range(3)]
    g3, s3 = np.stack([a for a, _ in stack]) > 0, np.stack([b for _, b in stack]) > 0
    assert np.allclose(eval_j(g3, s3), G['stack']['j'], rtol=0, atol=1e-12)
    assert np.allclose(eval_f(g3, s3), G['stack']['f'], rtol=0, atol=1e-12)


def test_evaluation_command(tmp_path):
    """Directory-level scoring: identical results score 1; the mean over pairs equals the mean of process_pair."""
    from click.testing import CliRunner
    import main
    from src.evaluation import evaluation_command_impl
    want = []
    for case in G['pairs'][:4]:
        seed = case['seed']
        gt, seg = mask_pair(seed, n_obj=1 + seed % 3, jitter=1 + seed % 4)
        for root, arr in (('gt', gt), ('res', seg), ('same', gt)):
            (tmp_path / root / 'clip').mkdir(parents=True, exist_ok=True)
            _save(tmp_path / root / 'clip' / f'{seed:05d}.png', arr)
        want.append(case['process_pair'])
    j, f, jf = evaluation_command_impl(str(tmp_path / 'gt'), str(tmp_path / 'res'), disable=True)
    w = np.array(want)
    assert np.isclose(j, w[:, 0].mean()) and np.isclose(f, w[:, 1].mean()) and np.isclose(jf, (j + f) / 2)
    assert evaluation_command_impl(str(tmp_path / 'gt'), str(tmp_path / 'same'), disable=True) == (1.0, 1.0, 1.0)
    r = CliRunner().invoke(main.cli, ['evaluation', '--help'])
    assert '--ground_truth' in r.output and '--computed_results' in r.output
