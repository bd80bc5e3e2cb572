"""Development aid: how evenly do the 148 CTAs of one fused launch finish?  Instrumented library needed:
  VOS_LIB_NAME=libvosprop_dbg.so VOS_NVCC_DEFS=-DVOS_KERNEL_DEBUG python semi-supervised-vos_b200/vosb200/build.py
Per CTA: nanoseconds between the start and the end of its MMA warp (globaltimer), for several target frames of one clip
(480p, R = 9, fp16, labels from the propagation itself), next to the number of segments (target tiles) of its range."""
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
for p in (str(REPO), str(REPO / 'semi-supervised-vos_b200')):
    sys.path.insert(0, p)
import ctypes as C  # noqa: E402

import torch  # noqa: E402

from vosb200 import PropagationEngine, synthetic  # noqa: E402
from vosb200 import _capi as capi  # noqa: E402
from vosb200.sequence import start_sequence  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    T = 40
    f, first = synthetic.clip_features(T, 480, 854, 3, seed=11, device=dev)
    f = f.half()
    P = 60 * 107
    eng = PropagationEngine(max_pixels=P, ring_slots=64, device=dev)
    start_sequence(eng, f[0], first, 4)
    eng.append_frames(1, f[1:T])
    buf = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    out = torch.empty((480, 854), dtype=torch.uint8, device=dev)
    durs, starts, own, smids = [], [], [], None
    for t in range(1, T):
        if t >= 30:
            capi.check(capi.lib().vosprop_debug_clocks(eng._h, buf.data_ptr()))
        eng.step(t, 40, 9, 8.0, 21.0, 1.0, False, want_prediction=False, want_lowres=False, want_fullres=False, out_fullres=out)
        if t >= 30:
            torch.cuda.synchronize()
            v = buf.view(148, 16).cpu()
            durs.append((v[:, 8] - v[:, 7].min()).double() / 1e3)        # end of each CTA relative to the first start (us)
            starts.append((v[:, 7] - v[:, 7].min()).double() / 1e3)      # start skew of each CTA (us)
            own.append(v[:, 6].double() / 1e3)                           # each CTA's own duration (us)
            smids = v[:, 9]
    d = torch.stack(durs)                                                   # (launches, 148)
    grid = C.c_int32()
    begins = (C.c_int64 * 149)()
    segs = C.c_int32()
    capi.check(capi.lib().vosprop_debug_decompose(P, 9, 148, C.byref(grid), begins, C.byref(segs)))
    tpf = (P + 127) // 128
    nt = 9 * tpf
    nseg = torch.tensor([(begins[c + 1] - 1) // nt - begins[c] // nt + 1 for c in range(148)])
    print('launch end per CTA (us): mean over CTAs %.1f, min %.1f, max %.1f (per launch: max - mean = %s)' %
          (d.mean(), d.min(), d.max(), ' '.join(f'{float(x):.1f}' for x in (d.max(1).values - d.mean(1)))))
    s_, o_ = torch.stack(starts), torch.stack(own)
    print('start skew per CTA (us): mean %.2f, max %.2f; own duration: mean %.1f, min %.1f, max %.1f, std over CTAs of the mean %.2f' %
          (s_.mean(), s_.max(), o_.mean(), o_.min(), o_.max(), o_.mean(0).std()))
    print('correlation over CTAs (means over launches): end vs start %.2f, end vs own duration %.2f' %
          (float(torch.corrcoef(torch.stack([d.mean(0), s_.mean(0)]))[0, 1]), float(torch.corrcoef(torch.stack([d.mean(0), o_.mean(0)]))[0, 1])))
    print('SM id of each CTA:', smids.tolist())
    print('mean own duration by CTA:', [round(float(x), 1) for x in o_.mean(0)])
    print('mean start by CTA:', [round(float(x), 1) for x in s_.mean(0)])
    m = d.mean(0)
    print('mean end by segments in the range: ' + ', '.join(f'{k} seg: {float(m[nseg == k].mean()):.1f} us ({int((nseg == k).sum())} CTAs)' for k in sorted(set(nseg.tolist()))))
    c = torch.corrcoef(d)          # correlation of the per-CTA pattern between launches
    print('correlation of the per-CTA end times between launches: mean off-diagonal %.2f' % float((c.sum() - c.diag().sum()) / (c.numel() - c.shape[0])))
    order = torch.argsort(m)
    print('fastest CTAs', [(int(i), round(float(m[i]), 1), int(nseg[i])) for i in order[:8]])
    print('slowest CTAs', [(int(i), round(float(m[i]), 1), int(nseg[i])) for i in order[-8:]])
    print('by CTA index (mean end, 8 per row):')
    for r in range(0, 148, 16):
        print('  ' + ' '.join(f'{float(x):6.1f}' for x in m[r:r + 16]))
    eng.close()


if __name__ == '__main__':
    main()
