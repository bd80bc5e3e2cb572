import sys, time, tempfile, os
from pathlib import Path
REPO = Path('/root/repo')
for p in (str(REPO), str(REPO / 'semi-supervised-vos_b200'), str(REPO / 'tools')):
    sys.path.insert(0, p)
import torch
import bench_cli
import src.inference as inf
from src.model.vos_net import VOSNet
import src.utils.inference_utils as iu

with tempfile.TemporaryDirectory() as tmp:
    root = Path(tmp)
    bench_cli.make_dataset(root, 8, 100)
    ckpt = root / 'ckpt.pth.tar'
    torch.manual_seed(0)
    torch.save({'state_dict': VOSNet('resnet50', pretrained=False).state_dict()}, ckpt)
    orig_load, orig_single = inf._load_net, inf.inference_single
    def timed_load(*a):
        t0 = time.perf_counter(); r = orig_load(*a); torch.cuda.synchronize(); print('  _load_net', round(time.perf_counter() - t0, 3)); return r
    def timed_single(*a, **k):
        t0 = time.perf_counter(); r = orig_single(*a, **k); print('  inference_single', round(time.perf_counter() - t0, 3)); return r
    inf._load_net, inf.inference_single = timed_load, timed_single
    orig_drain = iu._WRITER.drain
    def timed_drain():
        torch.cuda.synchronize(); t0 = time.perf_counter(); orig_drain(); print('  writer drain after GPU done', round(time.perf_counter() - t0, 3))
    iu._WRITER.drain = timed_drain
    for rep in range(3):
        t0 = time.perf_counter()
        inf.inference_command_impl(9, str(root), str(ckpt), 'resnet50', 1.0, 40, 8.0, 21.0, str(root / f'o{rep}'), 'cuda', 'single', None, 'resnet50', False, 1.15, 'mean', disable=True)
        print('total', round(time.perf_counter() - t0, 3))
