"""Development aid: does a second clip in flight (its own ClipSegmenter, engine and stream) raise the end-to-end rate?
The backbone of one clip can fill the SMs that idle in the tail of the other clip's fused launches and under its merge kernels.
    python tools/e2e_lanes_probe.py [n_sequences]"""
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
for p in (str(REPO), str(REPO / 'semi-supervised-vos_b200')):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from src.model.vos_net import VOSNet  # noqa: E402
from vosb200 import synthetic  # noqa: E402
from vosb200.pipeline import ClipSegmenter  # noqa: E402


def main():
    n_seq = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    dev = torch.device('cuda', 0)
    rs = np.random.RandomState(2017)
    lens = rs.randint(34, 105, size=n_seq)
    objs = rs.choice([1, 2, 3, 4], size=n_seq)
    torch.manual_seed(0)
    net = VOSNet('resnet50', pretrained=False)
    clips = [synthetic.clip_frames(int(lens[i]), 480, 854, int(objs[i]), seed=2000 + i, device=dev, raw=True) for i in range(n_seq)]
    outs = [torch.empty((int(lens[i]) - 1, 480, 854), dtype=torch.uint8, pin_memory=True) for i in range(n_seq)]
    frames = int(lens.sum()) - n_seq
    ref = None
    for lanes in (1, 2, 1, 2, 3):
        segs = [ClipSegmenter(net, device=dev) for _ in range(lanes)]
        streams = [torch.cuda.Stream(dev) for _ in range(lanes)]

        def step():
            for i, (f, first) in enumerate(clips):
                with torch.cuda.stream(streams[i % lanes]):
                    segs[i % lanes].segment(f, first, out=outs[i], sync=False)
            torch.cuda.synchronize(dev)

        step()
        if ref is None:
            ref = [o.clone() for o in outs]
        same = all(torch.equal(a, b) for a, b in zip(ref, outs))
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(2):
            step()
        dt = (time.perf_counter() - t0) / 2
        print(f'lanes {lanes}: {frames / dt:8.1f} frames/s end to end ({dt * 1e3:.1f} ms per pass over {n_seq} sequences), masks identical to one lane: {same}', flush=True)
        for s in segs:
            if s.engine is not None:
                s.engine.close()


if __name__ == '__main__':
    main()
