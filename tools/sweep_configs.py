"""Timing sweep over BASELINE.json's secondary configurations on one GPU (device time per propagated frame,
CUDA events around the engine's kernels; steady state, frame_idx >= 44 so both sigma branches are active):

  config 3: ref_num in {3, 5, 9, 12, 16, 20} x top-k in {full softmax, 5, 20, 50} at 480p
  config 4: 1080p (135 x 240 features), ref_num 9
  config 5: 10 objects (d = 11) at 480p

Prints one JSON object per line; `python tools/sweep_configs.py > profiles/<name>.jsonl`."""
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
for p in (str(REPO), str(REPO / 'semi-supervised-vos_b200')):
    sys.path.insert(0, p)
import torch  # noqa: E402

from vosb200 import PREC_F16, PREC_SPLIT3, PropagationEngine, plan_refs, synthetic  # noqa: E402
from vosb200.sequence import lowres_dims  # noqa: E402

K = 256


def measure(H, W, n_obj, ref_num, topk, prec, frames=8, t0=46, probability=False, warm=3, proto_scale=0.3, noise=0.1,
            block_skip=False, field_len=0.0):
    dev = torch.device('cuda', 0)
    T = t0 + warm + frames
    H_d, W_d = lowres_dims(H, W)
    P = H_d * W_d
    g = torch.Generator(device=dev).manual_seed(1)
    proto = torch.randn(n_obj + 1, K, device=dev, generator=g) * proto_scale
    eng = PropagationEngine(max_pixels=P, ring_slots=max(48, ref_num + 2), device=dev)
    eng.reset(H_d, W_d, H, W, n_obj + 1, prec)
    eng.block_skip('on' if block_skip is True else ('off' if block_skip is False else block_skip))
    cm = synthetic._class_map(synthetic._tracks(n_obj, torch.Generator().manual_seed(2)), 0, H_d, W_d, dev).reshape(-1)
    feat_dtype = torch.float16 if prec == PREC_F16 else torch.float32

    field = None
    if field_len > 0:
        # appearance that varies over the image like a texture: random Fourier features of the pixel position with
        # correlation length `field_len` (stride-8 pixels) and |f|^2 = 256 -- two pixels further apart than ~1.5 lengths
        # have nearly orthogonal embeddings, as different surfaces have in a trained network
        ys, xs = torch.meshgrid(torch.arange(H_d, device=dev, dtype=torch.float32), torch.arange(W_d, device=dev, dtype=torch.float32), indexing='ij')
        pos = torch.stack([ys.reshape(-1), xs.reshape(-1)], 1)
        w = torch.randn(2, K, device=dev, generator=g) / field_len
        b = torch.rand(K, device=dev, generator=g) * 6.2831853
        field = torch.cos(pos @ w + b) * (2.0 / K) ** 0.5 * 16.0

    def feature(t):
        if field is not None:
            f = field + proto[cm] + noise * torch.randn(P, K, device=dev, generator=g)
        else:
            f = proto[cm] + noise * torch.randn(P, K, device=dev, generator=g)
        return f.t().reshape(K, H_d, W_d).to(feat_dtype).contiguous()

    onehot = torch.zeros(n_obj + 1, P, device=dev).scatter_(0, cm.view(1, -1), 1.0)
    for t in range(t0 - 45, t0):                       # fill the ring's look-back window
        eng.append(t, feature(t))
        if probability:
            eng.set_labels_dense(t, onehot)
        else:
            eng.set_labels_index(t, cm.to(torch.uint8))
    feats = [feature(t) for t in range(t0, T)]
    out = torch.empty((H, W), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    for i, t in enumerate(range(t0, T)):
        if i == warm:                                  # the first frames warm clocks and caches up, untimed
            eng.enable_timing(8 * frames)
        eng.append(t, feats[i])
        refs, sig = plan_refs(t, 40, ref_num, 8.0, 21.0, probability)
        eng.propagate(t, refs, sig, 1.0, probability, want_prediction=False, want_lowres=False, want_fullres=False,
                      out_fullres=out, topk=topk)
    tm = eng.read_timing()
    eng.close()
    us = {k: v[0] / max(v[1], 1) * 1e3 for k, v in tm.items()}
    total = sum(us.values())
    flops = 2.0 * P * (len(refs) * P) * K
    return {'H': H, 'W': W, 'pixels': P, 'objects': n_obj, 'ref_num': ref_num, 'refs_used': len(refs), 'topk': topk,
            'precision': 'f16' if prec == PREC_F16 else 'split3', 'probability_propagation': probability, 'append_us': round(us['append'], 1),
            'affinity_us': round(us['affinity'], 1), 'merge_us': round(us['merge'], 1),
            'frames_per_s_device': round(1e6 / total, 1), 'affinity_tflops_algorithmic': round(flops / us['affinity'] / 1e6, 1)}


def main():
    if len(sys.argv) > 1 and sys.argv[1] == 'peaked':
        # embedding norm of trained features (|f|^2 ~ 256) instead of the low-contrast bench clips (|f|^2 ~ 26): blocks of the
        # affinity matrix that underflow to exactly zero are skipped by the fused kernel
        for ps, nz, fl in ((0.3, 0.1, 0.0), (0.95, 0.3, 0.0), (0.3, 0.1, 12.0), (0.3, 0.1, 6.0), (0.3, 0.1, 3.0)):
            for skip in (False, True, 'auto'):
                r = measure(480, 854, 2, 9, 0, PREC_F16, proto_scale=ps, noise=nz, block_skip=skip, field_len=fl)
                print(json.dumps(dict(config='peaked', class_prototype_norm2=round(K * ps * ps, 1), noise_norm2=round(K * nz * nz, 1),
                                      texture_field='none' if not fl else f'|f|^2 = 256, correlation length {fl} px', block_skip=skip, **r)), flush=True)
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'topk1':      # one short top-k run (k = 20) for profiling
        print(json.dumps(dict(config=3, **measure(480, 854, 2, 9, 20, PREC_F16, frames=4, warm=1))), flush=True)
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'prob1':      # one short probability-propagation run (dense-label kernel) for profiling
        print(json.dumps(dict(config='prob', **measure(480, 854, 2, 9, 0, PREC_F16, probability=True, frames=4, warm=1))), flush=True)
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'topk':
        for topk in (0, 5, 20, 50):
            print(json.dumps(dict(config=3, **measure(480, 854, 2, 9, topk, PREC_F16))), flush=True)
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'prob':
        for prec in (PREC_F16, PREC_SPLIT3):
            print(json.dumps(dict(config='prob', **measure(480, 854, 2, 9, 0, prec, probability=True))), flush=True)
        return
    for ref_num in (3, 5, 9, 12, 16, 20):
        for topk in (0, 5, 20, 50):
            print(json.dumps(dict(config=3, **measure(480, 854, 2, ref_num, topk, PREC_F16))), flush=True)
    print(json.dumps(dict(config=3, **measure(480, 854, 2, 9, 0, PREC_SPLIT3))), flush=True)
    print(json.dumps(dict(config=5, **measure(480, 854, 10, 9, 0, PREC_F16))), flush=True)
    for prec in (PREC_F16, PREC_SPLIT3):
        print(json.dumps(dict(config='prob', **measure(480, 854, 2, 9, 0, prec, probability=True))), flush=True)
    # the long 1080p kernels run into the power cap; they go last so the 480p lines above are not measured on a hot GPU
    for topk in (0, 20):
        print(json.dumps(dict(config=4, **measure(1080, 1920, 3, 9, topk, PREC_F16, frames=3, warm=1))), flush=True)
    print(json.dumps(dict(config=4, **measure(1080, 1920, 3, 9, 0, PREC_SPLIT3, frames=3, warm=1))), flush=True)

if __name__ == '__main__':
    main()
