"""Golden vectors for the J / F measures and the `evaluation` command's pair scoring, produced by the
reference's own ``src/utils/metrics.py`` and ``src/evaluation.py::process_pair`` on seeded masks.

Run in the build container (needs /root/reference):  python oracle/make_golden_eval.py
scikit-image is absent: the two primitives the F-measure takes from it (disk footprint, grey dilation) come from the scipy
shim in oracle/reference_harness.py -- J is pinned by the reference alone, F by the reference's code over that shim."""
import json
import sys
import tempfile
from pathlib import Path

import numpy as np
from PIL import Image

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from oracle import reference_harness as RH          # noqa: E402
from oracle.fixtures import mask_pair               # noqa: E402

GOLDEN = REPO / 'tests' / 'golden'


def palette():
    return RH.default_palette()


def main():
    ref = RH.import_evaluation(RH.import_reference('cpu'))
    out = {'pairs': [], 'note': 'F uses the scipy shim for skimage.morphology.disk / dilation'}
    for seed in range(8):
        gt, seg = mask_pair(seed, n_obj=1 + seed % 3, jitter=1 + seed % 4)
        per_obj = []
        for k in range(int(gt.max()) + 1):
            j, f = ref.metrics.evaluate_segmentation(gt == k, seg == k)
            per_obj.append([float(j), float(f)])
        with tempfile.TemporaryDirectory() as td:
            for name, arr in (('gt', gt), ('seg', seg)):
                img = Image.fromarray(arr, mode='P')
                img.putpalette(palette())
                img.save(Path(td) / f'{name}.png')
            pair = ref.evaluation.process_pair(Path(td) / 'gt.png', Path(td) / 'seg.png')
        out['pairs'].append({'seed': seed, 'per_object': per_obj, 'process_pair': [float(v) for v in pair]})
        print(seed, per_obj, pair)
    # corner cases of the measure (metrics.py:41-45, 104-113): empty masks
    z, o = np.zeros((40, 60), bool), np.zeros((40, 60), bool)
    o[10:20, 10:30] = True
    out['empty'] = {name: [float(v) for v in ref.metrics.evaluate_segmentation(a, b)]
                    for name, (a, b) in {'both_empty': (z, z), 'seg_empty': (o, z), 'gt_empty': (z, o)}.items()}
    # a (T,H,W) stack through eval_j / eval_f
    stack = [mask_pair(100 + t) for t in range(3)]
    g3, s3 = np.stack([a for a, _ in stack]) > 0, np.stack([b for _, b in stack]) > 0
    out['stack'] = {'j': [float(v) for v in ref.metrics.eval_j(g3, s3)], 'f': [float(v) for v in ref.metrics.eval_f(g3, s3)]}
    (GOLDEN / 'eval_jf.json').write_text(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
