"""Generate tests/golden/e2e_480p.npz by EXECUTING THE REFERENCE's `inference_command_impl` on CPU, from image frames
(BASELINE.json configs[0]: one synthetic 480p clip, 10 frames, 2 objects, random-init ResNet-50 with calibrated BatchNorm
statistics, ref_num 9, frame_range 40, sigma 8 / 21, temperature 1).

Run in the build container (needs /root/reference):   python oracle/make_golden_e2e.py
Stored: the masks the reference writes (read back from its PNGs), its per-frame predictions (d, P) (captured by wrapping
its predict() -- the sources are not touched), the embedding of frame 9 at 64 pixels, and sha256 of the inputs.
The frames and the checkpoint are rebuilt by the test from oracle/fixtures.py (numpy / name-seeded: same bytes everywhere).
"""
from __future__ import annotations

import hashlib
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch
from PIL import Image

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

from oracle import fixtures as FX  # noqa: E402
from oracle import reference_harness as RH  # noqa: E402

OUT = REPO / 'tests' / 'golden' / 'e2e_480p.npz'


def main():
    ref = RH.import_reference('cpu')
    cfg = FX.E2E
    frames, first = FX.e2e_frames(cfg['T'], cfg['H'], cfg['W'], cfg['n_objects'], cfg['seed'])
    torch.manual_seed(0)
    net = ref.vos_net.VOSNet('resnet50')
    state = FX.e2e_calibrated_state(net)
    preds, feats = [], {}
    probe = np.arange(64) * 100 + 7                 # 64 of the 6420 stride-8 pixels
    orig_predict = ref.inference_utils.predict

    def spy_predict(ref_feats, target, *a, **k):
        out = orig_predict(ref_feats, target, *a, **k)
        preds.append(out.detach().clone())
        if len(preds) == 9:
            feats[9] = target.detach().reshape(target.shape[0], -1)[:, probe].clone()
        return out

    ref.inference_utils.predict = spy_predict
    sys.path.insert(0, str(ref.root))
    try:
        import importlib
        inference = importlib.import_module('src.inference')
    finally:
        sys.path.remove(str(ref.root))
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        FX.e2e_write_tree(td / 'data', frames, first, cfg['video'])
        torch.save({'state_dict': state}, td / 'ckpt.pth')
        inference.inference_command_impl(9, str(td / 'data'), str(td / 'ckpt.pth'), 'resnet50', 1.0, 40, 8.0, 21.0, str(td / 'out'),
                                         'cpu', 'single', None, 'resnet50', False, 1.15, 'mean', disable=True)
        masks = np.stack([np.array(Image.open(td / 'out' / cfg['video'] / f'{t:05d}.png')) for t in range(1, cfg['T'])])
        jpeg_sha = hashlib.sha256(b''.join((td / 'data/JPEGImages/480p' / cfg['video'] / f'{t:05d}.jpg').read_bytes()
                                           for t in range(cfg['T']))).hexdigest()
    ref.inference_utils.predict = orig_predict
    assert len(preds) == cfg['T'] - 1
    state_sha = hashlib.sha256(b''.join(state[k].numpy().tobytes() for k in sorted(state))).hexdigest()
    np.savez_compressed(OUT, masks=masks.astype(np.uint8), preds=torch.stack(preds).numpy().astype(np.float32),
                        feat9_probe=feats[9].numpy().astype(np.float32), probe=probe, jpeg_sha256=jpeg_sha, state_sha256=state_sha)
    frac = [(masks == k).mean() for k in range(cfg['n_objects'] + 1)]
    print(f'wrote {OUT} ({OUT.stat().st_size / 1024:.0f} KB): masks {masks.shape}, class fractions {np.round(frac, 3)}, '
          f'first-frame fractions {[round(float((first == k).mean()), 3) for k in range(cfg["n_objects"] + 1)]}')


if __name__ == '__main__':
    main()
