"""Generate tests/golden/tta_*.npz by EXECUTING THE REFERENCE's test-time-augmentation loops
(src/utils/inference_utils.py:90-511) on seeded table-lookup embeddings.

Run in the build container (needs /root/reference):   python oracle/make_golden_tta.py
Stream A sees the seeded clip; stream B sees a second seeded embedding set (for the flip strategies: the
mirrored embeddings plus noise; for 2-scale: a clip generated at the second input size).
"""
from __future__ import annotations

import json
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

from oracle import propagation_oracle as O  # noqa: E402
from oracle import reference_harness as RH  # noqa: E402

OUT = REPO / 'tests' / 'golden'
SCALE = 1.15
#        name              strategy       T   H    W   objects seed prob   reduction
CASES = [('hor_flip',      'hor-flip',    8,  96, 160, 2,      71,  False, 'mean'),
         ('ver_flip',      'vert-flip',   8,  96, 160, 2,      72,  False, 'mean'),
         ('hor_flip_wide', 'hor-flip',    7, 128, 288, 3,      73,  False, 'mean'),
         ('hor_flip_prob', 'hor-flip',    7,  96, 160, 2,      74,  True,  'mean'),
         ('two_scale',     '2-scale',     7, 160, 320, 2,      75,  False, 'mean'),
         ('hor_two_scale', 'hor-2-scale', 7, 160, 320, 2,      76,  False, 'mean'),
         ('two_scale_prob', '2-scale',    6, 160, 320, 2,      77,  True,  'maximum'),
         ('multimodel',    'multimodel',  7,  96, 160, 2,      78,  False, 'mean')]


def streams(strategy, T, H, W, n_obj, seed):
    """(feats_a, feats_b, first annotation, size of input B)"""
    feats_a, first = O.synthetic_sequence(T, H, W, n_obj, seed=seed, feat_scale=0.30)
    g = torch.Generator().manual_seed(seed + 1000)
    if strategy in ('hor-flip', 'vert-flip'):
        flipped = torch.flip(feats_a, dims=(3,) if strategy == 'hor-flip' else (2,))
        return feats_a, flipped + 0.05 * torch.randn(flipped.shape, generator=g), first, None
    if strategy == 'multimodel':
        return feats_a, feats_a + 0.08 * torch.randn(feats_a.shape, generator=g), first, None
    Hb, Wb = int(np.ceil(H * SCALE)), int(np.ceil(W * SCALE))
    feats_b, _ = O.synthetic_sequence(T, Hb, Wb, n_obj, seed=seed, feat_scale=0.30)
    if strategy == 'hor-2-scale':
        feats_b = torch.flip(feats_b, dims=(3,))
    h2, w2 = int(np.ceil(H * O.SCALE * SCALE)), int(np.ceil(W * O.SCALE * SCALE))
    assert tuple(feats_b.shape[2:]) == (h2, w2), 'pick a size where the reference\'s two low-res computations agree'
    return feats_a, feats_b, first, (Hb, Wb)


def main():
    ref = RH.import_reference('cpu')
    meta = {}
    for name, strategy, T, H, W, n_obj, seed, prob, reduction in CASES:
        feats_a, feats_b, first, size_b = streams(strategy, T, H, W, n_obj, seed)
        with tempfile.TemporaryDirectory() as td:
            masks = RH.run_inference_two_streams(ref, strategy, feats_a, feats_b, first, RH.default_palette(), td,
                                                 size_b=size_b, probability_propagation=prob, reduction=reduction,
                                                 scale=SCALE)
        np.savez_compressed(OUT / f'tta_{name}.npz', masks=masks)
        meta[name] = dict(strategy=strategy, T=T, H=H, W=W, n_objects=n_obj, seed=seed, probability_propagation=prob,
                          reduction=reduction, scale=SCALE)
        print(name, 'live', [int((masks[-1] == c).sum()) for c in range(int(first.max()) + 1)])
    (OUT / 'meta_tta.json').write_text(json.dumps(meta, indent=1, sort_keys=True))


if __name__ == '__main__':
    main()
