"""Development aid, the other direction of tools/coresidency_probe.py: small blocks are already running on the SMs when
the fused kernel is launched on another stream.  Does it start next to them (ends ~its own duration later) or wait for them?"""
import ctypes as C
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
for p in (str(REPO), str(REPO / 'semi-supervised-vos_b200')):
    sys.path.insert(0, p)
import torch  # noqa: E402

from vosb200 import PREC_F16, PropagationEngine, plan_refs, synthetic  # noqa: E402


def main():
    lib = C.CDLL(str(REPO / 'tools' / 'probe' / 'libprobe.so'))
    dev = torch.device('cuda', 0)
    T = 20
    feats, first = synthetic.clip_features(T, 480, 854, 2, seed=1, device=dev)
    P = feats.shape[2] * feats.shape[3]
    f = feats.half()
    eng = PropagationEngine(max_pixels=P, device=dev)
    eng.reset(60, 107, 480, 854, 3, PREC_F16)
    for t in range(T):
        eng.append(t, f[t])
        eng.set_labels_index(t, torch.zeros(P, dtype=torch.uint8, device=dev))
    refs, sig = plan_refs(T - 1, 40, 9, 8.0, 21.0, False)
    kw = dict(write_labels=False, want_prediction=False, want_lowres=False, want_fullres=False)
    buf = torch.zeros(1 << 20, device=dev)
    sa, sb = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    for _ in range(20):
        eng.propagate(T - 1, refs, sig, **kw)
    torch.cuda.synchronize()
    for block, grid, iters in ((128, 51, 30000), (128, 148, 30000), (32, 201, 30000), (128, 296, 30000), (256, 148, 30000)):
        res = []
        for rep in range(4):
            e0, e_probe, e_aff = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            torch.cuda.synchronize()
            e0.record()
            with torch.cuda.stream(sb):
                sb.wait_event(e0)
                lib.probe_launch(grid, block, iters, C.c_void_p(buf.data_ptr()), C.c_void_p(sb.cuda_stream))
                e_probe.record(sb)
            with torch.cuda.stream(sa):
                sa.wait_event(e0)
                eng.propagate(T - 1, refs, sig, **kw)
                e_aff.record(sa)
            torch.cuda.synchronize()
            res.append((e0.elapsed_time(e_probe) * 1e3, e0.elapsed_time(e_aff) * 1e3))
        print(f'probe {grid} x {block} threads: probe ends at ' + ' '.join(f'{a:.0f}' for a, _ in res) +
              ' us; affinity + merge end at ' + ' '.join(f'{b:.0f}' for _, b in res) + ' us', flush=True)
    eng.close()


if __name__ == '__main__':
    main()
