"""Where does the end-to-end time go?  (diagnostic, run on the GPU box)"""
import sys, time
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO), str(REPO / 'semi-supervised-vos_b200')]
import torch
from src.model.vos_net import VOSNet

dev = torch.device('cuda')
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
net = VOSNet('resnet50', pretrained=False).to(dev).eval()

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n

for B in (35,):
    x = torch.randn(B, 3, 480, 854, device=dev)
    with torch.no_grad():
        t = timeit(lambda: net(x)); print(f'fp32 nchw            B={B:2d}: {t/B*1e3:.3f} ms/frame')
        with torch.autocast('cuda', dtype=torch.float16):
            t = timeit(lambda: net(x)); print(f'autocast fp16 nchw   B={B:2d}: {t/B*1e3:.3f} ms/frame')
        netcl = net.to(memory_format=torch.channels_last); xcl = x.contiguous(memory_format=torch.channels_last)
        with torch.autocast('cuda', dtype=torch.float16):
            t = timeit(lambda: netcl(xcl)); print(f'autocast fp16 nhwc   B={B:2d}: {t/B*1e3:.3f} ms/frame')
        import copy
        nh = copy.deepcopy(net).half().to(memory_format=torch.channels_last); xh = xcl.half()
        t = timeit(lambda: nh(xh)); print(f'half() nhwc          B={B:2d}: {t/B*1e3:.3f} ms/frame  ({166.9/(t/B*1e3):.0f} TFLOP/s)')
from vosb200.fused_backbone import FusedVOSNet
fused = FusedVOSNet(net)
for B in (10, 35):
    x = torch.randn(B, 3, 480, 854, device=dev)
    t = timeit(lambda: fused(x)); print(f'cuDNN-fused fp16 nhwc B={B:2d}: {t/B*1e3:.3f} ms/frame  ({166.9/(t/B*1e3):.0f} TFLOP/s)')
xp = torch.randn(35, 3, 480, 854).pin_memory()
t = timeit(lambda: xp.to(dev, non_blocking=True)); print(f'H2D 35 frames pinned: {t/35*1e3:.3f} ms/frame ({xp.numel()*4/t/1e9:.1f} GB/s)')
