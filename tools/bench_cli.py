"""Wall-clock frames/s of the drop-in entry point itself -- `inference_command_impl` (the body of `main.py inference`,
reference src/inference.py:54-113) -- on a synthetic DAVIS-shaped tree on disk: JPEG decode in DataLoader workers, VOSNet
on cuDNN under autocast, the propagation engine, palette PNGs written by the background writer.  Row N3 of SURVEY.md 8(f):
what a user of the command sees, as opposed to bench.py's in-memory numbers.  One JSON object per strategy."""
import json
import sys
import tempfile
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
for p in (str(REPO), str(REPO / 'semi-supervised-vos_b200')):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from PIL import Image  # noqa: E402


def make_dataset(root, videos=4, frames=50, H=480, W=854, seed=0):
    rs = np.random.RandomState(seed)
    palette = [0, 0, 0, 128, 0, 0, 0, 128, 0, 128, 128, 0] + [0] * (768 - 12)
    for v in range(videos):
        name = f'video{v}'
        (root / 'JPEGImages' / '480p' / name).mkdir(parents=True)
        (root / 'Annotations' / '480p' / name).mkdir(parents=True)
        base = np.asarray(Image.fromarray(rs.randint(0, 256, (H // 32 + 1, W // 32 + 1, 3)).astype(np.uint8)).resize((W, H), Image.BILINEAR))
        ann = np.zeros((H, W), np.uint8)
        for k in range(1, 3 + v % 2):
            y, x = rs.randint(40, H - 200), rs.randint(40, W - 300)
            ann[y:y + 150, x:x + 220] = k
        a = Image.fromarray(ann, mode='P')
        a.putpalette(palette)
        a.save(root / 'Annotations' / '480p' / name / '00000.png')
        for t in range(frames):
            img = np.roll(base, (3 * t, 5 * t), axis=(0, 1))
            Image.fromarray(img).save(root / 'JPEGImages' / '480p' / name / f'{t:05d}.jpg', quality=90)


def main():
    from src.inference import inference_command_impl
    from src.model.vos_net import VOSNet
    videos, frames = 8, 100
    with tempfile.TemporaryDirectory() as tmp:
        root = Path(tmp)
        make_dataset(root, videos, frames)
        ckpt = root / 'ckpt.pth.tar'
        torch.manual_seed(0)
        torch.save({'state_dict': VOSNet('resnet50', pretrained=False).state_dict()}, ckpt)
        # first 'single' warms cuDNN / the page cache up; `python tools/bench_cli.py single single` runs just those
        for strategy in (sys.argv[1:] or ('single', 'single', 'hor-flip', '3-scale')):
            save = root / f'out_{strategy}_{time.time_ns()}'
            t0 = time.perf_counter()
            inference_command_impl(9, str(root), str(ckpt), 'resnet50', 1.0, 40, 8.0, 21.0, str(save), 'cuda', strategy,
                                   None, 'resnet50', False, 1.15, 'mean', disable=True)
            dt = time.perf_counter() - t0
            n_png = len(list(save.glob('*/*.png')))
            print(json.dumps({'config': 'main.py inference (wall clock, model load and dataset read included)',
                              'strategy': strategy, 'gpu_jpeg': __import__('os').environ.get('VOS_GPU_JPEG', '1') != '0',
                              'host_cpus': len(__import__('os').sched_getaffinity(0)), 'videos': videos, 'frames': videos * frames, 'pngs_written': n_png,
                              'seconds': round(dt, 3), 'frames_per_s': round(videos * frames / dt, 1)}), flush=True)


if __name__ == '__main__':
    main()
