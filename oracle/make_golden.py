"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE's own functions on seeded inputs.

Run in the build container (needs /root/reference):   python oracle/make_golden.py
The reference has no tests or golden vectors of its own, so these files are
the parity pin: tests/test_oracle_golden.py checks the oracle restatement against them
(-m "not gpu") and tests/test_gpu_parity.py checks the CUDA path against them (-m gpu).
Inputs are regenerated from seeds at test time (oracle.synthetic_sequence); each file stores an
input checksum so a drifting generator is caught.
"""
from __future__ import annotations

import hashlib
import json
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

from oracle import propagation_oracle as O  # noqa: E402
from oracle import reference_harness as RH  # noqa: E402
from oracle.fixtures import seeded_state_dict  # noqa: E402

OUT = REPO / 'tests' / 'golden'

# name -> generator + propagation parameters.  Kept small: the whole golden set is < 1 MB.
SEQUENCES = {
    'A_label_r9': dict(T=20, H=96, W=160, n_objects=2, seed=11, feat_scale=0.30,
                       ref_num=9, frame_range=40, sigma_1=8.0, sigma_2=21.0, temperature=1.0,
                       probability_propagation=False),
    'B_prob_r5': dict(T=20, H=96, W=160, n_objects=3, seed=12, feat_scale=0.30,
                      ref_num=5, frame_range=6, sigma_1=8.0, sigma_2=21.0, temperature=1.0,
                      probability_propagation=True),
    'C_odd_temp': dict(T=18, H=100, W=150, n_objects=1, seed=13, feat_scale=0.35,
                       ref_num=4, frame_range=40, sigma_1=5.0, sigma_2=12.0, temperature=0.5,
                       probability_propagation=False),
    'D_long_wrap': dict(T=60, H=64, W=80, n_objects=2, seed=14, feat_scale=0.30,
                        ref_num=9, frame_range=40, sigma_1=8.0, sigma_2=21.0, temperature=1.0,
                        probability_propagation=False),
    'E_many_objects': dict(T=12, H=72, W=120, n_objects=9, seed=15, feat_scale=0.30,
                           ref_num=9, frame_range=40, sigma_1=8.0, sigma_2=21.0,
                           temperature=1.0, probability_propagation=False),
    # W_d >= 32: exercised by the index-label tensor-core kernel (vos_affinity_idx)
    'F_wide_r9': dict(T=20, H=128, W=288, n_objects=3, seed=16, feat_scale=0.30,
                      ref_num=9, frame_range=40, sigma_1=8.0, sigma_2=21.0, temperature=1.0,
                      probability_propagation=False),
    'G_wide_odd_r5': dict(T=24, H=200, W=264, n_objects=5, seed=17, feat_scale=0.30,
                          ref_num=5, frame_range=10, sigma_1=6.0, sigma_2=15.0, temperature=1.3,
                          probability_propagation=False),
}
GEN_KEYS = ('T', 'H', 'W', 'n_objects', 'seed', 'feat_scale')


def checksum(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    ref = RH.import_reference('cpu')
    torch.set_grad_enabled(False)

    # ---- P1 sample_frames (src/model/predict.py:74-89)
    rows = []
    for ref_num in (3, 4, 5, 9, 12, 20):
        for rng in (6, 10, 40):
            for t in range(1, 130):
                idx = ref.predict.sample_frames(t, rng, ref_num).tolist()
                rows.append([ref_num, rng, t, len(idx)] + idx + [-1] * (20 - len(idx)))
    np.savez_compressed(OUT / 'sample_frames.npz', table=np.asarray(rows, dtype=np.int32))

    # ---- P2 get_spatial_weight (src/model/predict.py:158-175)
    sw = {}
    for (h, w, s) in ((9, 13, 8.0), (9, 13, 21.0), (5, 7, 3.0), (12, 20, 8.0)):
        sw[f'w_{h}_{w}_{s:g}'] = ref.predict.get_spatial_weight((h, w), s).numpy()
    np.savez_compressed(OUT / 'spatial_weight.npz', **sw)

    # ---- P4/P7 get_labels / index_to_onehot  (predict.py:92-96, utils.py:59-68)
    lab = {}
    for (H, W, nobj, seed) in ((96, 160, 2, 11), (100, 150, 1, 13), (61, 83, 4, 3), (480, 854, 3, 5)):
        _, first = O.synthetic_sequence(1, H, W, nobj, K=8, seed=seed)
        d = int(first.max()) + 1
        H_d, W_d = O.lowres_dims(H, W)
        one = ref.predict.get_labels(torch.from_numpy(first.astype(np.int64)), d, H, W, H_d, W_d)
        lab[f'first_{H}_{W}_{nobj}_{seed}'] = first
        lab[f'low_{H}_{W}_{nobj}_{seed}'] = one[:, 0].argmax(0).numpy().astype(np.uint8)
        assert (one.sum(0) == 1).all()
    np.savez_compressed(OUT / 'first_frame_labels.npz', **lab)

    # ---- P3/P5/P6/P8: the real inference_single on table-lookup features
    meta = {}
    for name, cfg in SEQUENCES.items():
        gen = {k: cfg[k] for k in GEN_KEYS}
        feats, first = O.synthetic_sequence(gen['T'], gen['H'], gen['W'], gen['n_objects'],
                                            seed=gen['seed'], feat_scale=gen['feat_scale'])
        run = {k: v for k, v in cfg.items() if k not in GEN_KEYS}
        with tempfile.TemporaryDirectory() as td:
            masks, preds = RH.run_inference_single(ref, feats, first, RH.default_palette(), td, **run)
        preds = torch.stack(preds).numpy()
        np.savez_compressed(OUT / f'seq_{name}.npz', masks=masks, predictions=preds)
        # logit statistics of the last frame, to document that the fixture is not degenerate
        P = feats.shape[2] * feats.shape[3]
        S = feats[-2].permute(1, 2, 0).reshape(P, -1) @ feats[-1].reshape(feats.shape[1], P)
        top2 = S.topk(2, dim=0).values
        live = [int((masks[-1] == c).sum()) for c in range(int(first.max()) + 1)]
        meta[name] = dict(cfg, features_sha256=checksum(feats),
                          first_sha256=hashlib.sha256(first.tobytes()).hexdigest(),
                          logit_std=float(S.std()), top1_top2_gap_median=float((top2[0] - top2[1]).median()),
                          last_mask_class_pixels=live)
        print(name, 'logit std %.2f' % meta[name]['logit_std'],
              'gap %.2f' % meta[name]['top1_top2_gap_median'], 'live', live)

    # ---- P3 stand-alone predict() calls incl. frame_idx regimes, duplicated refs, int32 labels
    cases = {}
    feats, first = O.synthetic_sequence(50, 72, 104, 2, seed=21, feat_scale=0.30)
    T, K, H_d, W_d = feats.shape
    P = H_d * W_d
    low, d = O.first_frame_labels(first)
    g = torch.Generator().manual_seed(5)
    hist = torch.stack([O.index_to_onehot(torch.randint(0, d, (P,), generator=g), d) for _ in range(T)], 1)
    hist[:, 0] = O.index_to_onehot(low, d)
    prob_hist = torch.rand(d, T, P, generator=g)
    prob_hist /= prob_hist.sum(0, keepdim=True)
    wd = ref.predict.get_spatial_weight((H_d, W_d), 8.0)
    ws = ref.predict.get_spatial_weight((H_d, W_d), 21.0)
    for (t, rn, rng, temp, prob) in ((1, 9, 40, 1.0, False), (4, 9, 40, 1.0, False), (9, 9, 40, 1.0, False),
                                     (10, 9, 40, 1.0, False), (15, 9, 40, 1.0, False), (16, 9, 40, 1.0, False),
                                     (17, 3, 40, 1.0, False), (17, 4, 40, 1.0, False), (30, 9, 40, 2.0, False),
                                     (49, 20, 40, 1.0, False), (49, 9, 5, 1.0, False), (12, 9, 40, 1.0, True),
                                     (40, 12, 40, 0.7, True)):
        lab_hist = prob_hist if prob else hist
        if t == 1 and not prob:
            lab_in = hist[:, :1].to(torch.int32)  # the reference's first call gets int32 labels (predict.py:96)
        else:
            lab_in = lab_hist[:, :t]
        out = ref.predict.predict(feats[:t], feats[t], lab_in, None if prob else wd, None if prob else ws,
                                  t, rng, rn, temp, prob)
        cases[f't{t}_n{rn}_r{rng}_T{temp:g}_p{int(prob)}'] = out.numpy()
    np.savez_compressed(OUT / 'predict_cases.npz', **cases)
    meta['predict_cases'] = dict(T=50, H=72, W=104, n_objects=2, seed=21, feat_scale=0.30, label_seed=5,
                                 features_sha256=checksum(feats), hist_sha256=checksum(hist),
                                 prob_hist_sha256=checksum(prob_hist))

    # ---- P0 VOSNet.forward (src/model/vos_net.py:42-51) with name-seeded weights
    torch.manual_seed(0)
    net = ref.vos_net.VOSNet('resnet50').eval()
    sd = seeded_state_dict(net.state_dict())
    net.load_state_dict(sd)
    x = torch.randn(1, 3, 64, 96, generator=torch.Generator().manual_seed(3))
    y = net(x)
    np.savez_compressed(OUT / 'vosnet_forward.npz', y=y.numpy())
    meta['vosnet_forward'] = dict(keys={k: list(v.shape) for k, v in sd.items()}, input_seed=3,
                                  input_shape=[1, 3, 64, 96])
    # ---- signatures of the drop-in callables
    import inspect
    sig = lambda f: list(inspect.signature(f).parameters)  # noqa: E731
    sys.path.insert(0, str(ref.root))
    try:
        import importlib
        inf = importlib.import_module('src.inference')
    finally:
        sys.path.remove(str(ref.root))
    meta['signatures'] = {
        'predict': sig(ref.predict.predict), 'sample_frames': sig(ref.predict.sample_frames),
        'prepare_first_frame': sig(ref.predict.prepare_first_frame), 'get_labels': sig(ref.predict.get_labels),
        'get_spatial_weight': sig(ref.predict.get_spatial_weight),
        'inference_single': sig(ref.inference_utils.inference_single),
        'inference_command_impl': sig(inf.inference_command_impl),
        'inference_command_options': [p.name for p in inf.inference_command.params]}
    (OUT / 'meta.json').write_text(json.dumps(meta, indent=1, sort_keys=True))
    print('wrote', sorted(p.name for p in OUT.iterdir()))


if __name__ == '__main__':
    main()
