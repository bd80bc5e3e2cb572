"""GPU parity of the drop-in entry points themselves (semi-supervised-vos_b200/src/utils/inference_utils.py):
`inference_single` and the two-stream test-time-augmentation loops, driven exactly like the reference's are driven by
oracle/reference_harness.py -- a table-lookup model, a one-clip loader, an annotation PNG on disk -- and compared
with the PNGs the REFERENCE wrote for the same inputs (tests/golden/seq_*.npz, tta_*.npz)."""
import numpy as np
import pytest
import torch
from PIL import Image

from oracle import reference_harness as RH
from tests import _golden as G

pytestmark = pytest.mark.gpu


def _dataset(tmp_path, first, video='clip'):
    ann_dir = tmp_path / 'Annotations' / '480p'
    (ann_dir / video).mkdir(parents=True)
    img = Image.fromarray(np.asarray(first).astype(np.uint8), mode='P')
    img.putpalette(RH.default_palette())
    img.save(ann_dir / video / '00000.png')
    return ann_dir, tmp_path / 'out'


def _read_masks(save, video, T):
    return np.stack([np.asarray(Image.open(save / video / f'{t:05d}.png')) for t in range(1, T)]).astype(np.uint8)


def _table(feats):
    """'Network' that looks the embedding of every frame of the batch up by the frame number encoded in the input."""
    return lambda inp: torch.stack([feats[int(v)] for v in inp[:, 0, 0, 0].tolist()])


@pytest.fixture()
def mirror():
    from src.config import Config
    from src.utils import inference_utils as iu
    old = Config.DEVICE
    Config.DEVICE = torch.device('cuda', 0)
    yield iu
    Config.DEVICE = old


@pytest.mark.parametrize('name', ['A_label_r9', 'B_prob_r5', 'F_wide_r9'])
def test_inference_single_writes_the_reference_pngs(name, tmp_path, mirror):
    feats, first, run = G.sequence_inputs(name)
    masks_ref, _ = G.sequence_golden(name)
    T, (H, W) = feats.shape[0], first.shape
    ann_dir, save = _dataset(tmp_path, first)
    loader = [(torch.full((1, 1, H, W), float(t)), ('clip',)) for t in range(T)]
    with torch.no_grad():
        mirror.inference_single(_table(feats.cuda()), loader, T, ann_dir, 'clip', str(save), run['sigma_1'], run['sigma_2'],
                                run['frame_range'], run['ref_num'], run['temperature'], run['probability_propagation'], True)
    masks = _read_masks(save, 'clip', T)
    agree = float((masks == masks_ref).mean())
    print(f'inference_single/{name}: mask agreement {agree:.6f}')
    assert agree >= 0.999
    assert np.array_equal(np.asarray(Image.open(save / 'clip' / '00000.png')), first)      # predict.py:120-126


@pytest.mark.parametrize('name', G.TTA_NAMES)
def test_two_stream_strategies_write_the_reference_pngs(name, tmp_path, mirror):
    cfg, feats_a, feats_b, first, size_b = G.tta_inputs(name)
    want = np.load(G.GOLDEN / f'tta_{name}.npz')['masks']
    T, (H, W) = cfg['T'], first.shape
    Hb, Wb = (H, W) if size_b is None else size_b
    ann_dir, save = _dataset(tmp_path, first)
    fa, fb = feats_a.cuda(), feats_b.cuda()
    common = (T, ann_dir, 'clip', str(save), 8.0, 21.0, 40, 9, 1.0, cfg['probability_propagation'])
    strategy = cfg['strategy']
    with torch.no_grad():
        if strategy == 'multimodel':
            loader = [(torch.full((1, 1, H, W), float(t)), ('clip',)) for t in range(T)]
            mirror.inference_multimodel(_table(fa), _table(fb), loader, *common, cfg['reduction'], True)
        else:
            loader = [([torch.full((1, 1, H, W), float(t)), torch.full((1, 1, Hb, Wb), float(t) + 0.25)], ('clip',))
                      for t in range(T)]
            model = lambda inp: torch.stack([(fb if v % 1 else fa)[int(v)] for v in inp[:, 0, 0, 0].tolist()])  # noqa: E731
            if strategy == 'hor-flip':
                mirror.inference_hor_flip(model, loader, *common, cfg['reduction'], True)
            elif strategy == 'vert-flip':
                mirror.inference_ver_flip(model, loader, *common, cfg['reduction'], True)
            else:
                mirror.inference_2_scale(model, loader, *common, cfg['scale'], cfg['reduction'], strategy == 'hor-2-scale', True)
    masks = _read_masks(save, 'clip', T)
    agree = float((masks == want).mean())
    print(f'{strategy}/{name}: mask agreement {agree:.6f}')
    assert agree >= 0.999


@pytest.mark.parametrize('name', ['three_scale', 'three_scale_prob'])
def test_three_scale_writes_the_reference_pngs(name, tmp_path, mirror):
    """inference_3_scale against the PNGs the reference's own inference_3_scale wrote (oracle/make_golden_3scale.py):
    two videos, three passes over the loader, 480 x 910 outputs, element-wise maximum of the class indices."""
    import json
    from oracle import propagation_oracle as O
    cfg = json.loads((G.GOLDEN / 'meta_3scale.json').read_text())[name]
    want = np.load(G.GOLDEN / f'tta_{name}.npz')
    H, W, scale = cfg['H'], cfg['W'], cfg['scale']
    feats, by_shape, videos = {}, {}, [v['name'] for v in cfg['videos']]
    ann_dir = None
    for v in cfg['videos']:
        per_scale = []
        for k, s in enumerate((0.9, 1.0, scale)):
            Hs, Ws = int(np.ceil(H * s)), int(np.ceil(W * s))
            f, lab = O.synthetic_sequence(v['T'], Hs, Ws, v['objects'], seed=v['seed'], feat_scale=0.30)
            per_scale.append(f.cuda())
            by_shape[(Hs, Ws)] = k
            if s == 1.0:
                first = lab
        feats[v['name']] = per_scale
        (tmp_path / 'Annotations' / '480p' / v['name']).mkdir(parents=True)
        img = Image.fromarray(first.astype(np.uint8), mode='P')
        img.putpalette(RH.default_palette())
        img.save(tmp_path / 'Annotations' / '480p' / v['name'] / '00000.png')
    ann_dir, save = tmp_path / 'Annotations' / '480p', tmp_path / 'out'

    def model(inp):
        k = by_shape[tuple(inp.shape[2:])]
        return torch.stack([feats[videos[int(c) // 1000]][k][int(c) % 1000] for c in inp[:, 0, 0, 0].tolist()])

    loader = [(torch.full((1, 1, H, W), float(1000 * vi + t)), (v['name'],)) for vi, v in enumerate(cfg['videos'])
              for t in range(v['T'])]
    with torch.no_grad():
        mirror.inference_3_scale(model, loader, len(loader), ann_dir, videos[0], str(save), 8.0, 21.0, 40, 9, 1.0,
                                 cfg['probability_propagation'], scale, True)
    for v in cfg['videos']:
        masks = _read_masks(save, v['name'], v['T'])
        assert masks.shape == want[v['name']].shape == (v['T'] - 1, 480, 910)
        agree = float((masks == want[v['name']]).mean())
        print(f"3-scale/{name}/{v['name']}: mask agreement {agree:.6f}")
        assert agree >= 0.999
