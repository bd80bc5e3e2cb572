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


# ------------------------------------------------------------------------------------------------------------------
# End-to-end fixture of BASELINE.json configs[0]: a synthetic 480p clip as a JPEG tree + a random-init
# ResNet-50 VOSNet whose BatchNorm statistics are calibrated by four seeded train-mode forwards (with the statistics of a
# fresh module -- mean 0, variance 1 -- a random network maps every frame to nearly the same embedding).
# ------------------------------------------------------------------------------------------------------------------
E2E = dict(T=10, H=480, W=854, n_objects=2, seed=2024, video='clip0')


def e2e_frames(T=10, H=480, W=854, n_objects=2, seed=2024):
    """(frames (T,H,W,3) uint8, first annotation (H,W) uint8): textured background and textured objects (ellipses) that
    drift a few pixels per frame.  numpy only: the same bytes on every machine."""
    rs = np.random.RandomState(seed)
    yy, xx = np.mgrid[:H, :W].astype(np.float32)

    def texture(n_waves):
        t = np.zeros((H, W, 3), np.float32)
        for _ in range(n_waves):
            fy, fx = rs.uniform(-0.08, 0.08, 2)
            ph = rs.uniform(0, 6.283, 3)
            amp = rs.uniform(0.05, 0.18, 3)
            t += amp * np.sin(yy[..., None] * fy + xx[..., None] * fx + ph)
        return t

    base = rs.uniform(0.25, 0.75, (n_objects + 1, 3)).astype(np.float32)
    tex = [texture(6) for _ in range(n_objects + 1)]
    cy, cx = rs.uniform(0.3, 0.7, n_objects) * H, rs.uniform(0.25, 0.75, n_objects) * W
    ry, rx = rs.uniform(0.12, 0.2, n_objects) * H, rs.uniform(0.08, 0.15, n_objects) * W
    vy, vx = rs.uniform(-4, 4, n_objects), rs.uniform(-6, 6, n_objects)
    frames = np.zeros((T, H, W, 3), np.uint8)
    first = None
    for t in range(T):
        cls = np.zeros((H, W), np.uint8)
        for k in range(n_objects):
            cls[((yy - cy[k] - vy[k] * t) / ry[k]) ** 2 + ((xx - cx[k] - vx[k] * t) / rx[k]) ** 2 <= 1.0] = k + 1
        img = np.zeros((H, W, 3), np.float32)
        for k in range(n_objects + 1):
            sel = cls == k
            # object textures move with their object
            dy, dx = (0, 0) if k == 0 else (int(round(vy[k - 1] * t)), int(round(vx[k - 1] * t)))
            img[sel] = (base[k] + np.roll(tex[k], (dy, dx), (0, 1)))[sel]
        img += rs.normal(0, 0.01, img.shape).astype(np.float32)
        frames[t] = (np.clip(img, 0, 1) * 255).round().astype(np.uint8)
        if t == 0:
            first = cls
    return frames, first


def e2e_write_tree(root, frames, first, video='clip0'):
    """DAVIS layout: <root>/JPEGImages/480p/<video>/%05d.jpg + <root>/Annotations/480p/<video>/00000.png (palette PNG)."""
    from pathlib import Path

    from PIL import Image
    root = Path(root)
    (root / 'JPEGImages/480p' / video).mkdir(parents=True, exist_ok=True)
    (root / 'Annotations/480p' / video).mkdir(parents=True, exist_ok=True)
    for t, f in enumerate(frames):
        Image.fromarray(f).save(root / 'JPEGImages/480p' / video / f'{t:05d}.jpg', quality=95)
    ann = Image.fromarray(first, mode='P')
    ann.putpalette([0, 0, 0, 128, 0, 0, 0, 128, 0, 128, 128, 0] + [0] * (256 * 3 - 12))
    ann.save(root / 'Annotations/480p' / video / '00000.png')


def e2e_calibrated_state(net, seed=0, n_forwards=4, size=(2, 3, 240, 432)):
    """Random initialisation (seeded, torch's default init of `net`'s constructor is replaced by name-seeded weights) and
    BatchNorm statistics from `n_forwards` train-mode forwards on seeded frames of the fixture's statistics.  `net`: a
    VOSNet('resnet50') of the reference or of this repository (same state-dict keys).  Returns the state dict."""
    net.load_state_dict(seeded_state_dict(net.state_dict()))
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.reset_running_stats()
            m.momentum = None                         # cumulative average over the calibration forwards
    frames, _ = e2e_frames(T=n_forwards * size[0], H=size[2], W=size[3], seed=seed + 77)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    x = (torch.from_numpy(frames).permute(0, 3, 1, 2).float() / 255.0 - mean) / std
    net.train()
    with torch.no_grad():
        for i in range(n_forwards):
            net(x[i * size[0]:(i + 1) * size[0]])
    net.eval()
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.momentum = 0.1
    return {k: v.clone() for k, v in net.state_dict().items()}
