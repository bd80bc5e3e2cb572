"""Development aid: time one 480p propagation step (R = 9, fp16 embeddings) with parts of the fused epilogue
switched off (vosprop_debug_flags) to see which part bounds the kernel.  Results with flags != 0 are wrong."""
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
for p in (str(REPO), str(REPO / 'semi-supervised-vos_b200')):
    sys.path.insert(0, p)
import torch  # noqa: E402

from vosb200 import PREC_F16, PREC_SPLIT3, PropagationEngine, plan_refs, synthetic  # noqa: E402
from vosb200 import _capi as capi  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    T = 20
    feats, first = synthetic.clip_features(T, 480, 854, 2, seed=1, device=dev)
    P = feats.shape[2] * feats.shape[3]
    for prec, f in ((PREC_F16, feats.half()),):
        eng = PropagationEngine(max_pixels=P, device=dev)
        eng.reset(60, 107, 480, 854, 3, prec)
        g = torch.Generator(device=dev).manual_seed(0)
        for t in range(T):
            eng.append(t, f[t])
            lab = (torch.rand(P, device=dev, generator=g) < 0.03).to(torch.uint8) if t else torch.zeros(P, dtype=torch.uint8, device=dev)
            lab[: P // 3] = 1
            eng.set_labels_index(t, lab)
        refs, sig = plan_refs(T - 1, 40, 9, 8.0, 21.0, False)
        for flags in (0,):
            capi.check(capi.lib().vosprop_debug_flags(eng._h, flags))
            for _ in range(3):
                eng.propagate(T - 1, refs, sig, write_labels=False, want_prediction=False, want_lowres=False, want_fullres=False)
            eng.enable_timing(64)
            for _ in range(20):
                eng.propagate(T - 1, refs, sig, write_labels=False, want_prediction=False, want_lowres=False, want_fullres=False)
            tm = eng.read_timing()
            eng.enable_timing(0)
            clk = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
            capi.check(capi.lib().vosprop_debug_clocks(eng._h, clk.data_ptr()))
            eng.propagate(T - 1, refs, sig, write_labels=False, want_prediction=False, want_lowres=False, want_fullres=False)
            torch.cuda.synchronize()
            capi.check(capi.lib().vosprop_debug_clocks(eng._h, None))
            c = clk.view(148, 16).double().mean(0).tolist()
            cv = clk.view(148, 16)
            print('   cycles/CTA  producer: total %.0f wait_empty %.0f | mma: total %.0f wait_q %.0f wait_acc_empty %.0f wait_full %.0f | '
                  'ns/CTA %.0f  clock %.3f GHz  kernel span %.1f us'
                  % (tuple(c[:6]) + (c[6], c[2] / c[6], (int(cv[:, 8].max()) - int(cv[:, 7].min())) / 1e3)))
            print(f'prec={prec} flags={flags:2d}: affinity {tm["affinity"][0] / tm["affinity"][1] * 1e3:7.1f} us   merge {tm["merge"][0] / tm["merge"][1] * 1e3:5.1f} us', flush=True)
        eng.close()


if __name__ == '__main__':
    main()
