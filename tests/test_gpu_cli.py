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
    from src.model.vos_net import VOSNet
    from src.utils.datasets import InferenceDataset
    root, ckpt, firsts, sd = env
    save = root / 'out'
    inference_command_impl(9, str(root), str(ckpt), 'resnet50', 1.0, 40, 8.0, 21.0, str(save), 'cuda', 'single', None,
                           'resnet50', False, 1.15, 'mean', disable=True)
    net = VOSNet('resnet50', pretrained=False)
    net.load_state_dict(sd)
    net = net.cuda().eval()
    ds = InferenceDataset(str(root / 'JPEGImages' / '480p'), disable=True)
    feats = {}
    with torch.no_grad(), torch.autocast('cuda', dtype=torch.float16):
        for i in range(len(ds)):
            img, video = ds[i]
            feats.setdefault(video, []).append(net(img[None].cuda())[0].float().cpu())
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
