"""Development aid: host time of one PropagationEngine.step()/propagate() call against the device time of the step
(480p, R = 9, fp16).  If the host needs as long per call as the device per step, the per-frame loop is host-bound."""
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
for p in (str(REPO), str(REPO / 'semi-supervised-vos_b200')):
    sys.path.insert(0, p)
import torch  # noqa: E402

from vosb200 import PREC_F16, PropagationEngine, plan_refs, synthetic  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    T, N = 20, 400
    feats, first = synthetic.clip_features(T, 480, 854, 2, seed=1, device=dev)
    P = feats.shape[2] * feats.shape[3]
    f = feats.half()
    eng = PropagationEngine(max_pixels=P, device=dev)
    eng.reset(60, 107, 480, 854, 3, PREC_F16)
    for t in range(T):
        eng.append(t, f[t])
        eng.set_labels_index(t, torch.zeros(P, dtype=torch.uint8, device=dev))
    refs, sig = plan_refs(T - 1, 40, 9, 8.0, 21.0, False)
    full = torch.empty((480, 854), dtype=torch.uint8, device=dev)
    kw = dict(write_labels=True, want_prediction=False, want_lowres=False, want_fullres=False, out_fullres=full)
    for _ in range(50):
        eng.propagate(T - 1, refs, sig, **kw)
    torch.cuda.synchronize()
    for name, fn in (('propagate()', lambda: eng.propagate(T - 1, refs, sig, **kw)),
                     ('step() = plan_refs + propagate', lambda: eng.step(T - 1, 40, 9, 8.0, 21.0, 1.0, False, **kw))):
        for rep in range(3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for _ in range(N):
                fn()
            e1.record()
            t_host = (time.perf_counter() - t0) / N * 1e6
            torch.cuda.synchronize()
            print(f'{name}: host {t_host:.1f} us per call (loop issue time), device {e0.elapsed_time(e1) / N * 1e3:.1f} us per step', flush=True)
    eng.close()


if __name__ == '__main__':
    main()
