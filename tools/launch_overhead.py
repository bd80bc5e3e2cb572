"""Development aid: how much of an event-bracketed launch of the fused kernel is the kernel?  One 480p step (R = 9, fp16):
(a) CUDA events around every launch (what bench.py's roofline uses), (b) one event pair around N back-to-back steps."""
import sys
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
    g = torch.Generator(device=dev).manual_seed(0)
    for t in range(T):
        eng.append(t, f[t])
        lab = (torch.rand(P, device=dev, generator=g) < 0.03).to(torch.uint8) if t else torch.zeros(P, dtype=torch.uint8, device=dev)
        lab[: P // 3] = 1
        eng.set_labels_index(t, lab)
    refs, sig = plan_refs(T - 1, 40, 9, 8.0, 21.0, False)
    kw = dict(write_labels=False, want_prediction=False, want_lowres=False, want_fullres=False)
    for _ in range(20):
        eng.propagate(T - 1, refs, sig, **kw)
    torch.cuda.synchronize()
    eng.enable_timing(N + 8)
    for _ in range(N):
        eng.propagate(T - 1, refs, sig, **kw)
    tm = eng.read_timing()
    eng.enable_timing(0)
    a = tm['affinity'][0] / tm['affinity'][1] * 1e3
    m = tm['merge'][0] / tm['merge'][1] * 1e3
    def chain(append=False, merge=True, **kws):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(N):
            if append:
                eng.append(T - 1, f[T - 1])
            eng.propagate(T - 1, refs, sig, **kws)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / N * 1e3

    print(f'per-launch events: affinity {a:.1f} us  merge {m:.1f} us', flush=True)
    full = torch.empty((480, 854), dtype=torch.uint8, device=dev)
    kw2 = dict(kw, write_labels=True)
    variants = [('affinity + merge (no outputs)', dict(kw)),
                ('affinity + merge + ring labels', kw2),
                ('affinity + merge + full-resolution mask', dict(kw, out_fullres=full)),
                ('affinity + merge + labels + mask (the clip loop)', dict(kw2, out_fullres=full)),
                ('append + affinity + merge + labels + mask (round-2 loop)', dict(kw2, out_fullres=full, append=True))]
    for _ in range(3):          # warm clocks
        chain(**variants[0][1])
    res = {name: [] for name, _ in variants}
    for rep in range(4):        # interleaved repeats: clock / power drift shows as spread, not as a difference between variants
        for name, kws in variants:
            res[name].append(chain(**kws))
    for name, v in res.items():
        print(f'  chain {name}: ' + ' '.join(f'{x:.1f}' for x in v) + ' us per step', flush=True)
    eng.close()


if __name__ == '__main__':
    main()
