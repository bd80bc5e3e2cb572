"""Development aid: do small blocks co-reside with a resident vos_affinity_idx CTA?  Stream A runs one 480p affinity launch
(R = 9); stream B starts a probe kernel (tools/probe/probe.cu, <= 32 registers) right after the affinity kernel has started.
If the probe finishes within a few microseconds it ran next to the affinity CTAs; if it takes ~the kernel time it queued."""
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
    prio = -1 if 'prio' in sys.argv else 0
    sa, sb = torch.cuda.Stream(dev), torch.cuda.Stream(dev, priority=prio)
    print('probe stream priority', prio, flush=True)
    for _ in range(20):
        eng.propagate(T - 1, refs, sig, **kw)
    torch.cuda.synchronize()
    for block in (32, 128):
        for grid in (148, 592):
            res = []
            for rep in range(5):
                e_start, e_aff = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e_go, e_probe = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                with torch.cuda.stream(sa):
                    # a first launch keeps the GPU busy while the host queues the rest; the probe waits for its end
                    eng.propagate(T - 1, refs, sig, **kw)
                    e_start.record(sa)
                    eng.propagate(T - 1, refs, sig, **kw)
                    e_aff.record(sa)
                with torch.cuda.stream(sb):
                    sb.wait_event(e_start)
                    e_go.record(sb)
                    lib.probe_launch(grid, block, 2000, C.c_void_p(buf.data_ptr()), C.c_void_p(sb.cuda_stream))
                    e_probe.record(sb)
                torch.cuda.synchronize()
                res.append((e_go.elapsed_time(e_probe) * 1e3, e_start.elapsed_time(e_aff) * 1e3))
            print(f'block {block:4d} grid {grid:5d}: probe done after ' + ' '.join(f'{a:.0f}' for a, _ in res) +
                  ' us; affinity + merge took ' + ' '.join(f'{b:.0f}' for _, b in res) + ' us', flush=True)
    # the probe alone
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    lib.probe_launch(592, 32, 2000, C.c_void_p(buf.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    e1.record()
    torch.cuda.synchronize()
    print(f'probe alone (592 x 32): {e0.elapsed_time(e1) * 1e3:.0f} us')
    eng.close()


if __name__ == '__main__':
    main()
