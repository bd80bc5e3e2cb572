"""Shared helpers: load tests/golden fixtures and regenerate their seeded inputs."""
import hashlib
import json
from pathlib import Path

import numpy as np
import torch

from oracle import propagation_oracle as O

GOLDEN = Path(__file__).resolve().parent / 'golden'
META = json.loads((GOLDEN / 'meta.json').read_text())
GEN_KEYS = ('T', 'H', 'W', 'n_objects', 'seed', 'feat_scale')
SEQ_NAMES = [k for k in META if k[0] in 'ABCDEFG' and k[1] == '_']
META16 = json.loads((GOLDEN / 'meta16.json').read_text())   # oracle/make_golden_f16.py
SEQ16_NAMES = sorted(META16)
DTYPES = {'float16': torch.float16, 'bfloat16': torch.bfloat16}


def sha(t):
    a = t.contiguous().numpy() if isinstance(t, torch.Tensor) else t
    return hashlib.sha256(a.tobytes()).hexdigest()


def sequence_inputs(name):
    cfg = META[name]
    feats, first = O.synthetic_sequence(cfg['T'], cfg['H'], cfg['W'], cfg['n_objects'],
                                        seed=cfg['seed'], feat_scale=cfg['feat_scale'])
    assert sha(feats) == cfg['features_sha256'], 'seeded input generator drifted from the golden'
    run = {k: v for k, v in cfg.items() if k in ('ref_num', 'frame_range', 'sigma_1', 'sigma_2',
                                                 'temperature', 'probability_propagation')}
    return feats, first, run


def sequence16_inputs(tag):
    """(16-bit embeddings, first annotation, run parameters): the seeded clip rounded once to fp16 / bf16."""
    cfg = META16[tag]
    feats, first, run = sequence_inputs(cfg['source'])
    feats = feats.to(DTYPES[cfg['dtype']])
    assert sha(feats.float()) == cfg['features_sha256']
    return feats, first, run


def sequence16_golden(tag):
    z = np.load(GOLDEN / f'seq16_{tag}.npz')
    return z['masks'], z['predictions']


def sequence_golden(name):
    z = np.load(GOLDEN / f'seq_{name}.npz')
    return z['masks'], z['predictions']


def predict_case_inputs():
    cfg = META['predict_cases']
    feats, first = O.synthetic_sequence(cfg['T'], cfg['H'], cfg['W'], cfg['n_objects'],
                                        seed=cfg['seed'], feat_scale=cfg['feat_scale'])
    T, K, H_d, W_d = feats.shape
    P = H_d * W_d
    low, d = O.first_frame_labels(first)
    g = torch.Generator().manual_seed(cfg['label_seed'])
    hist = torch.stack([O.index_to_onehot(torch.randint(0, d, (P,), generator=g), d) for _ in range(T)], 1)
    hist[:, 0] = O.index_to_onehot(low, d)
    prob_hist = torch.rand(d, T, P, generator=g)
    prob_hist /= prob_hist.sum(0, keepdim=True)
    assert sha(feats) == cfg['features_sha256'] and sha(hist) == cfg['hist_sha256']
    assert sha(prob_hist) == cfg['prob_hist_sha256']
    return feats, hist, prob_hist


def parse_case(key):
    # 't16_n9_r40_T1_p0'
    parts = key.split('_')
    return dict(t=int(parts[0][1:]), ref_num=int(parts[1][1:]), frame_range=int(parts[2][1:]),
                temperature=float(parts[3][1:]), prob=bool(int(parts[4][1:])))


META_TTA = json.loads((GOLDEN / 'meta_tta.json').read_text())   # oracle/make_golden_tta.py
TTA_NAMES = sorted(META_TTA)


def tta_inputs(name):
    """(cfg, feats_a, feats_b, first annotation, (H,W) of input B) exactly as oracle/make_golden_tta.py builds them."""
    import sys
    sys.path.insert(0, str(GOLDEN.parent.parent))
    from oracle.make_golden_tta import streams
    cfg = META_TTA[name]
    feats_a, feats_b, first, size_b = streams(cfg['strategy'], cfg['T'], cfg['H'], cfg['W'], cfg['n_objects'], cfg['seed'])
    return cfg, feats_a, feats_b, first, size_b
