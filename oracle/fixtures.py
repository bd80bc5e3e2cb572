"""Seeded fixtures shared by the golden generator and the tests.  TEST INFRASTRUCTURE ONLY."""
import hashlib

import numpy as np
import torch


def seeded_state_dict(template: dict) -> dict:
    """Deterministic weights keyed by parameter name (so any module with the same keys/shapes
    gets the same values regardless of construction order)."""
    out = {}
    for k in sorted(template):
        v = template[k]
        g = torch.Generator().manual_seed(int(hashlib.sha256(k.encode()).hexdigest()[:8], 16))
        if k.endswith('num_batches_tracked'):
            out[k] = torch.zeros_like(v)
        elif k.endswith('running_var'):
            out[k] = torch.rand(v.shape, generator=g) + 0.5
        elif k.endswith('running_mean') or k.endswith('bias'):
            out[k] = torch.randn(v.shape, generator=g) * 0.1
        elif v.dim() == 1:
            out[k] = torch.rand(v.shape, generator=g) * 0.5 + 0.75
        else:
            fan_in = v[0].numel()
            out[k] = torch.randn(v.shape, generator=g) * (1.0 / fan_in) ** 0.5
    return out


def mask_pair(seed, H=120, W=200, n_obj=2, jitter=3):
    """A label map of blobs and a perturbed copy of it (shifted / eroded objects), like a propagated mask."""
    rs = np.random.RandomState(seed)
    gt = np.zeros((H, W), np.uint8)
    seg = np.zeros((H, W), np.uint8)
    yy, xx = np.mgrid[:H, :W]
    for k in range(1, n_obj + 1):
        cy, cx, ry, rx = rs.randint(20, H - 20), rs.randint(30, W - 30), rs.randint(10, 30), rs.randint(15, 45)
        gt[((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1] = k
        dy, dx, s = rs.randint(-jitter, jitter + 1), rs.randint(-jitter, jitter + 1), 1.0 + 0.1 * rs.randn()
        seg[((yy - cy - dy) / (ry * s)) ** 2 + ((xx - cx - dx) / (rx * s)) ** 2 <= 1] = k
    return gt, seg
