"""Seeded synthetic DAVIS-shaped clips for benchmarks and smoke runs (no dataset, no network).
Built with torch ops on the target device; nothing here is on the propagation path."""
from __future__ import annotations

from typing import Tuple

import torch

from .sequence import lowres_dims

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def _tracks(n_objects: int, g: torch.Generator):
    r = lambda lo, hi: torch.rand(n_objects, generator=g) * (hi - lo) + lo  # noqa: E731
    return dict(cy=r(.2, .8), cx=r(.2, .8), ry=r(.10, .22), rx=r(.10, .22), vy=r(-.012, .012), vx=r(-.012, .012))


def _class_map(tr, t: int, hh: int, ww: int, device) -> torch.Tensor:
    y = ((torch.arange(hh, device=device, dtype=torch.float32) + .5) / hh)[:, None]
    x = ((torch.arange(ww, device=device, dtype=torch.float32) + .5) / ww)[None, :]
    cm = torch.zeros(hh, ww, dtype=torch.long, device=device)
    for k in range(tr['cy'].numel()):
        inside = ((y - float(tr['cy'][k] + tr['vy'][k] * t)) / float(tr['ry'][k])) ** 2 + \
                 ((x - float(tr['cx'][k] + tr['vx'][k] * t)) / float(tr['rx'][k])) ** 2 <= 1.0
        cm = torch.where(inside, torch.full_like(cm, k + 1), cm)
    return cm


def clip_features(T: int, H: int, W: int, n_objects: int, seed: int, device, K: int = 256,
                  feat_scale: float = 0.30, noise: float = 0.35) -> Tuple[torch.Tensor, torch.Tensor]:
    """Embedding-space clip: (features (T,K,H_d,W_d) fp32 on `device`, first annotation (H,W) uint8).
    Each class owns a prototype embedding; objects are drifting ellipses; white noise on top."""
    g = torch.Generator().manual_seed(seed)
    H_d, W_d = lowres_dims(H, W)
    proto = (torch.randn(n_objects + 1, K, generator=g) * feat_scale).to(device)
    tr = _tracks(n_objects, g)
    gd = torch.Generator(device=device).manual_seed(seed + 1)
    feats = torch.empty(T, K, H_d, W_d, device=device)
    for t in range(T):
        f = proto[_class_map(tr, t, H_d, W_d, device)]
        f = f + noise * feat_scale * torch.randn(H_d, W_d, K, device=device, generator=gd)
        feats[t] = f.permute(2, 0, 1)
    return feats, _class_map(tr, 0, H, W, device).to(torch.uint8)


def clip_frames(T: int, H: int, W: int, n_objects: int, seed: int, device, pinned: bool = True, raw: bool = False
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Image-space clip in pinned host memory + first annotation (H,W) uint8 on host.  raw = False: (T,3,H,W) fp32
    ImageNet-normalised frames (what InferenceDataset hands the reference's loop); raw = True: (T,H,W,3) uint8 decoded RGB
    (what InferenceDataset(raw=True) hands this build's loop: normalisation runs on the GPU)."""
    g = torch.Generator().manual_seed(seed)
    tr = _tracks(n_objects, g)
    colors = torch.rand(n_objects + 1, 3, generator=g).to(device)
    gd = torch.Generator(device=device).manual_seed(seed + 1)
    mean = torch.tensor(IMAGENET_MEAN, device=device).view(3, 1, 1)
    std = torch.tensor(IMAGENET_STD, device=device).view(3, 1, 1)
    out = torch.empty((T, H, W, 3), dtype=torch.uint8, pin_memory=pinned) if raw else \
        torch.empty((T, 3, H, W), dtype=torch.float32, pin_memory=pinned)
    for t in range(T):
        img = colors[_class_map(tr, t, H, W, device)].permute(2, 0, 1)
        img = (img + 0.05 * torch.randn(3, H, W, device=device, generator=gd)).clamp_(0, 1)
        if raw:
            out[t].copy_((img * 255.0).round().to(torch.uint8).permute(1, 2, 0))
        else:
            out[t].copy_((img - mean) / std)
    return out, _class_map(tr, 0, H, W, device).to(torch.uint8).cpu()
