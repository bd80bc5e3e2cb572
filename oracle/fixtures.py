"""Seeded fixtures shared by the golden generator and the tests.  TEST INFRASTRUCTURE ONLY."""
import hashlib

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
