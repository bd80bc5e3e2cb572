"""The drop-in entry point end to end on the GPU: `inference_command_impl` (the body of `main.py inference`,
reference src/inference.py:54-113) on a DAVIS-layout dataset on disk -- JPEG decode, VOSNet on cuDNN under
autocast, the engine, palette PNGs -- for the `single` and `hor-flip` strategies.

The backbone runs in fp16 here and in fp32 in the CPU reference, so the check is two-fold: (1) the files the
reference writes exist with the right size, mode and palette, frame 0 is the annotation; (2) the masks equal what
the ORACLE propagates from the very embeddings this model produced for the same JPEGs (everything downstream of the
backbone is then covered by one comparison)."""
import numpy as np
import pytest
import torch
from PIL import Image

from oracle import propagation_oracle as O
from oracle import reference_harness as RH
from oracle.fixtures import seeded_state_dict

pytestmark = pytest.mark.gpu

H, W, T = 96, 160, 6


def _dataset(root):
    g = np.random.default_rng(5)
    firsts = {}
    for v, n_obj in (('bear', 2), ('cars', 1)):
        (root / 'JPEGImages' / '480p' / v).mkdir(parents=True)
        (root / 'Annotations' / '480p' / v).mkdir(parents=True)
        _, first = O.synthetic_sequence(1, H, W, n_obj, K=8, seed=len(v))
        firsts[v] = first
        ann = Image.fromarray(first, mode='P')
        ann.putpalette(RH.default_palette())
        ann.save(root / 'Annotations' / '480p' / v / '00000.png')
        base = g.integers(0, 255, (H, W, 3), dtype=np.uint8)
        for t in range(T):
            img = np.roll(base, (2 * t, 3 * t), axis=(0, 1)).copy()
            img[first > 0] = (img[first > 0] // 2 + np.array([120, 30, 30], dtype=np.uint8))
            Image.fromarray(img).save(root / 'JPEGImages' / '480p' / v / f'{t:05d}.jpg', quality=95)
    return firsts


@pytest.fixture()
def env(tmp_path):
    from src.config import Config
    from src.model.vos_net import VOSNet
    old = Config.DEVICE
    firsts = _dataset(tmp_path)
    net = VOSNet('resnet50', pretrained=False)
    sd = seeded_state_dict(net.state_dict())
    ckpt = tmp_path / 'ckpt.pth.tar'
    torch.save({'state_dict': sd}, ckpt)
    yield tmp_path, ckpt, firsts, sd
    Config.DEVICE = old


def test_inference_command_single_end_to_end(env):
    from src.inference import inference_command_impl
    from src.utils.datasets import InferenceDataset
    root, ckpt, firsts, sd = env
    save = root / 'out'
    inference_command_impl(9, str(root), str(ckpt), 'resnet50', 1.0, 40, 8.0, 21.0, str(save), 'cuda', 'single', None,
                           'resnet50', False, 1.15, 'mean', disable=True)
    # the same network object the command builds (cuDNN-fused inference form), fed the way the command feeds it: the
    # frames of one video, up to BACKBONE_LOOKAHEAD at a time
    import src.inference as inf
    from src.utils.inference_utils import BACKBONE_LOOKAHEAD
    net = inf._load_net('resnet50', str(ckpt))
    ds = InferenceDataset(str(root / 'JPEGImages' / '480p'), disable=True)
    frames = {}
    for i in range(len(ds)):
        img, video = ds[i]
        frames.setdefault(video, []).append(img)
    feats = {}
    with torch.no_grad(), torch.autocast('cuda', dtype=torch.float16):
        for video, imgs in frames.items():
            for b0 in range(0, len(imgs), BACKBONE_LOOKAHEAD):
                out = net(torch.stack(imgs[b0:b0 + BACKBONE_LOOKAHEAD]).cuda())
                feats.setdefault(video, []).extend(f.float().cpu() for f in out)
    for video, first in firsts.items():
        frame0 = Image.open(save / video / '00000.png')
        assert frame0.mode == 'P' and np.array_equal(np.asarray(frame0), first)
        want, _ = O.propagate_sequence(torch.stack(feats[video]), first)
        for t in range(1, T):
            png = Image.open(save / video / f'{t:05d}.png')
            assert png.mode == 'P' and png.size == (W, H) and png.getpalette()[:36] == RH.default_palette()[:36]
        got = np.stack([np.asarray(Image.open(save / video / f'{t:05d}.png')) for t in range(1, T)])
        agree = float((got == want.numpy()).mean())
        print(f'CLI single/{video}: mask agreement with the oracle on the same embeddings {agree:.6f}')
        assert agree >= 0.999


def test_inference_command_hor_flip_end_to_end(env):
    from src.inference import inference_command_impl
    root, ckpt, firsts, _ = env
    save = root / 'out_flip'
    inference_command_impl(9, str(root), str(ckpt), 'resnet50', 1.0, 40, 8.0, 21.0, str(save), 'cuda', 'hor-flip', None,
                           'resnet50', False, 1.15, 'mean', disable=True)
    for video in firsts:
        for t in range(T):
            png = Image.open(save / video / f'{t:05d}.png')
            assert png.mode == 'P' and png.size == (W, H)


def test_cpu_device_is_refused(env):
    from src.inference import inference_command_impl
    root, ckpt, _, _ = env
    with pytest.raises(RuntimeError, match='no CPU path'):
        inference_command_impl(9, str(root), str(ckpt), 'resnet50', 1.0, 40, 8.0, 21.0, str(root / 'o'), 'cpu', 'single',
                               None, 'resnet50', False, 1.15, 'mean', disable=True)


def test_gpu_normalisation_is_bit_identical_to_torchvision(env):
    """vosprop_normalize_u8 against the reference's host pipeline (ToTensor + Normalize, datasets.py:128-131,147): fp32
    output equal bit for bit on every byte value and on real frames (odd pixel counts included); fp16 = that value rounded;
    and the raw dataset mode delivers exactly the frames the normalised mode was computed from."""
    from torchvision import transforms
    from src.utils.datasets import InferenceDataset
    from vosb200 import normalize_frames
    host = transforms.Compose([transforms.ToTensor(), transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    g = torch.Generator().manual_seed(0)
    for shape in ((1, 16, 16, 3), (2, 37, 53, 3), (1, 1, 1, 3), (3, 96, 160, 3)):
        rgb = torch.randint(0, 256, shape, dtype=torch.uint8, generator=g)
        if shape[1] == 16:
            rgb.view(-1)[:768] = torch.arange(256, dtype=torch.uint8).repeat_interleave(3)     # every byte value in every channel
        want = torch.stack([host(f.numpy()) for f in rgb])
        got32 = normalize_frames(rgb.cuda(), torch.float32)
        assert got32.shape == want.shape and got32.is_contiguous(memory_format=torch.channels_last)
        assert torch.equal(got32.cpu(), want)
        assert torch.equal(normalize_frames(rgb.cuda(), torch.float16).cpu(), want.half())
    root = env[0]
    plain = InferenceDataset(str(root / 'JPEGImages' / '480p'), disable=True)
    raw = InferenceDataset(str(root / 'JPEGImages' / '480p'), disable=True, raw=True)
    for i in (0, 3, len(plain) - 1):
        (a, va), (b, vb) = plain[i], raw[i]
        assert va == vb and b.dtype == torch.uint8 and tuple(b.shape) == (H, W, 3)
        assert torch.equal(normalize_frames(b[None].cuda(), torch.float32)[0].cpu(), a)
    with pytest.raises(TypeError):
        normalize_frames(torch.zeros(1, 3, 8, 8, dtype=torch.uint8, device='cuda'))
