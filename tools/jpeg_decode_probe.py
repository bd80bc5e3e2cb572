"""Development aid: nvJPEG (torchvision.io.decode_jpeg on the GPU) against PIL on 480p frames -- throughput and pixel differences.
The drop-in keeps PIL in the loader because its pixels are the reference's; this measures what an opt-in GPU decode would give."""
import io
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
for p in (str(REPO), str(REPO / 'semi-supervised-vos_b200')):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from PIL import Image  # noqa: E402
from torchvision.io import decode_jpeg  # noqa: E402

from vosb200 import synthetic  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    frames, _ = synthetic.clip_frames(48, 480, 854, 3, seed=5, device=dev, raw=True)
    # photographic texture on top of the flat synthetic colours, so that the DCT has something to do
    g = torch.Generator().manual_seed(1)
    tex = torch.nn.functional.interpolate(torch.rand(48, 3, 60, 107, generator=g), size=(480, 854), mode='bilinear').permute(0, 2, 3, 1)
    imgs = (frames.float() * 0.6 + tex * 100).clamp(0, 255).to(torch.uint8).numpy()
    for quality, sub in ((90, 2), (75, 2), (95, 0)):
        blobs = []
        for a in imgs:
            b = io.BytesIO()
            Image.fromarray(a).save(b, format='JPEG', quality=quality, subsampling=sub)
            blobs.append(b.getvalue())
        t0 = time.perf_counter()
        pil = [np.asarray(Image.open(io.BytesIO(b)).convert('RGB')) for b in blobs]
        t_pil = (time.perf_counter() - t0) / len(blobs)
        datas = [torch.frombuffer(bytearray(b), dtype=torch.uint8) for b in blobs]
        out = decode_jpeg(datas, device=dev)           # warm-up (creates the nvJPEG handle)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            out = decode_jpeg(datas, device=dev)
        torch.cuda.synchronize()
        t_gpu = (time.perf_counter() - t0) / (3 * len(blobs))
        t0 = time.perf_counter()
        for _ in range(3):
            one = [decode_jpeg(d, device=dev) for d in datas]
        torch.cuda.synchronize()
        t_one = (time.perf_counter() - t0) / (3 * len(blobs))
        diff = torch.stack([(o.permute(1, 2, 0).cpu().int() - torch.from_numpy(p).int()).abs() for o, p in zip(out, pil)])
        print(f'quality {quality} subsampling {sub} ({np.mean([len(b) for b in blobs]) / 1e3:.0f} KB per frame): PIL {t_pil * 1e3:.2f} ms per frame on one '
              f'core; nvJPEG batched {t_gpu * 1e3:.3f} ms per frame ({1 / t_gpu:.0f} frames/s), one by one {t_one * 1e3:.3f} ms; '
              f'pixels that differ {float((diff > 0).float().mean()) * 100:.2f} %, by more than 1 level {float((diff > 1).float().mean()) * 100:.3f} %, '
              f'max {int(diff.max())}', flush=True)


if __name__ == '__main__':
    main()
