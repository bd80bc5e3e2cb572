"""Generate tests/golden/seq16_*.npz: the reference's own inference_single on FP16-VALUED embeddings.

Run in the build container (needs /root/reference):   python oracle/make_golden_f16.py
On CUDA the reference extracts embeddings under autocast (src/utils/inference_utils.py:35,52-53), so what
predict() sees there are fp16 values.  These goldens feed the reference (CPU, fp32 arithmetic) the seeded
synthetic embeddings rounded once to fp16 (or bf16) -- the inputs the single-pass tensor-core modes
(VOSPROP_PREC_F16 / _BF16) keep exactly.  Same sequences and parameters as make_golden.py.
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
from oracle.make_golden import GEN_KEYS, SEQUENCES, checksum  # noqa: E402

OUT = REPO / 'tests' / 'golden'
CASES = [('A_label_r9', torch.float16), ('B_prob_r5', torch.float16), ('D_long_wrap', torch.float16),
         ('E_many_objects', torch.float16), ('F_wide_r9', torch.float16), ('F_wide_r9', torch.bfloat16)]


def main():
    ref = RH.import_reference('cpu')
    torch.set_grad_enabled(False)
    meta = {}
    for name, dt in CASES:
        cfg = SEQUENCES[name]
        gen = {k: cfg[k] for k in GEN_KEYS}
        feats, first = O.synthetic_sequence(gen['T'], gen['H'], gen['W'], gen['n_objects'], seed=gen['seed'],
                                            feat_scale=gen['feat_scale'])
        feats = feats.to(dt).float()
        run = {k: v for k, v in cfg.items() if k not in GEN_KEYS}
        with tempfile.TemporaryDirectory() as td:
            masks, preds = RH.run_inference_single(ref, feats, first, RH.default_palette(), td, **run)
        tag = f"{name}_{'f16' if dt == torch.float16 else 'bf16'}"
        np.savez_compressed(OUT / f'seq16_{tag}.npz', masks=masks, predictions=torch.stack(preds).numpy())
        meta[tag] = dict(cfg, source=name, dtype=str(dt).replace('torch.', ''), features_sha256=checksum(feats))
        print(tag, 'live', [int((masks[-1] == c).sum()) for c in range(int(first.max()) + 1)])
    (OUT / 'meta16.json').write_text(json.dumps(meta, indent=1, sort_keys=True))


if __name__ == '__main__':
    main()
