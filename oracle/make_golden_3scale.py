"""Generate tests/golden/tta_three_scale*.npz by EXECUTING THE REFERENCE's ``inference_3_scale``
(src/utils/inference_utils.py:514-595) on seeded table-lookup embeddings: two videos, three passes over the loader.

Run in the build container (needs /root/reference):   python oracle/make_golden_3scale.py
Only the fused masks are stored (bit-packed would be smaller; npz compression is enough: they are piecewise constant)."""
import json
import sys
import tempfile
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from oracle import propagation_oracle as O  # noqa: E402
from oracle import reference_harness as RH  # noqa: E402

OUT = REPO / 'tests' / 'golden'
SCALE = 1.15
#        name                H    W   prob   videos: (name, T, objects, seed)
CASES = [('three_scale',      160, 320, False, (('bear', 6, 2, 81), ('cars', 5, 3, 82))),
         ('three_scale_prob', 160, 320, True,  (('bear', 5, 2, 83),))]


def clip(T, H, W, n_obj, seed):
    """Embeddings of one clip at the three input sizes + its first annotation (at the loader's size)."""
    feats, first = [], None
    for s in (0.9, 1.0, SCALE):
        Hs, Ws = int(np.ceil(H * s)), int(np.ceil(W * s))
        f, lab = O.synthetic_sequence(T, Hs, Ws, n_obj, seed=seed, feat_scale=0.30)
        want = (int(np.ceil(H * O.SCALE * s)), int(np.ceil(W * O.SCALE * s)))
        assert tuple(f.shape[2:]) == want, 'pick a size where the reference\'s two low-res computations agree'
        feats.append(f)
        if s == 1.0:
            first = lab
    return feats, first


def main():
    ref = RH.import_reference('cpu')
    meta = {}
    for name, H, W, prob, videos in CASES:
        feats = {v: clip(T, H, W, n, seed)[0] for v, T, n, seed in videos}
        firsts = {v: clip(T, H, W, n, seed)[1] for v, T, n, seed in videos}
        with tempfile.TemporaryDirectory() as td:
            got = RH.run_inference_3_scale(ref, feats, firsts, RH.default_palette(), td, (H, W), scale=SCALE,
                                           probability_propagation=prob)
        for v in got:
            want = O.propagate_three_scales(feats[v], firsts[v], SCALE, probability_propagation=prob).numpy()
            agree = float((want == got[v]).mean())
            print(name, v, got[v].shape, 'oracle agreement', agree, 'classes', np.unique(got[v]).tolist())
            assert agree == 1.0
        np.savez_compressed(OUT / f'tta_{name}.npz', **{v: got[v] for v in got})
        meta[name] = dict(H=H, W=W, probability_propagation=prob, scale=SCALE,
                          videos=[dict(name=v, T=T, objects=n, seed=seed) for v, T, n, seed in videos])
    (OUT / 'meta_3scale.json').write_text(json.dumps(meta, indent=1))


if __name__ == '__main__':
    main()
