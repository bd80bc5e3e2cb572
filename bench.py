#!/usr/bin/env python
"""bench.py -- propagated frames/s at 480p (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): a synthetic DAVIS-2017-val-shaped shard per GPU -- `clips` clips x
`frames` frames at 480x854 (60x107 features, K=256), 2-4 objects, ref_num 9, frame_range 40,
sigma 8/21, temperature 1 (BASELINE.json configs[1]: 30 seqs x ~70 frames over 8 GPUs = ~4 clips
per GPU).  A *step* is one pass over the shard.  Weak scaling: every rank owns its own shard, no
collective on the hot path; NCCL only gathers the per-sequence results at the end.

  value : propagation-stage throughput (ring append + fused affinity/softmax/prior/gather kernel +
          merge/write-back), stride-8 embeddings already resident in HBM, timed with CUDA events.
  e2e   : the same metric through the public API (vosb200.pipeline.ClipSegmenter.segment): frames in
          pinned host memory -> H2D -> VOSNet on cuDNN -> propagation -> uint8 masks -> D2H, per step.
  roofline     : the fused affinity kernel against the measured bf16 tensor peak (algorithmic FLOPs
                 2*P*(R*P)*K per launch, CUDA events around every launch of the timed region).
  cpu_baseline : the CPU port of the reference (oracle/) on this box's host cores, bounded sample.
--impl reference times that CPU port end to end (the reference is pure Python/torch and cannot
travel to the GPU box; the oracle is its line-by-line restatement, pinned bit-exact to it).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
for _p in (str(REPO), str(REPO / 'semi-supervised-vos_b200')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

H, W, K = 480, 854, 256
REF_NUM, FRAME_RANGE, SIGMA_1, SIGMA_2, TEMPERATURE = 9, 40, 8.0, 21.0, 1.0
METRIC = 'propagated frames/sec at 480p'
# dram__bytes_read.sum + dram__bytes_write.sum of one vos_affinity_idx launch (R = 9) from the ncu --set full
# captures under profiles/ (see profiles/README.md); None until captured for that mode
TRAFFIC_BYTES = {'f16': 33.7e6, 'split3': 63.4e6}
UNIT = 'frames/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', choices=['ours', 'reference'], default='ours')
    ap.add_argument('--clips', type=int, default=4, help='clips per GPU per step')
    ap.add_argument('--frames', type=int, default=70, help='frames per clip')
    ap.add_argument('--precision', choices=['f16', 'split3'], default='f16',
                    help='embeddings resident in HBM for `value`: fp16 (what VOSNet emits under autocast; one exact '
                         'tensor-core pass) or fp32 (bf16 hi+lo split, three passes)')
    ap.add_argument('--lanes', type=int, default=1,
                    help='sequences in flight per GPU (one engine + stream each; 2 gives +2-3 %% frames/s but the lanes\' small '
                         'kernels delay the start of the other lane\'s fused kernel, so its per-launch time reads higher)')
    ap.add_argument('--no-kernel-events', action='store_true', help='do not bracket every kernel with CUDA events (roofline fields become null)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--ref-frames', type=int, default=2, help='reference arm: propagated frames per step')
    return ap.parse_args()


def workload_config(args, n_gpus):
    return {'workload': f'synthetic DAVIS-2017-val-shaped shard: {args.clips} clips x {args.frames} frames per GPU, '
                        f'480x854 (60x107 stride-8 features, K=256), 2-4 objects',
            'clips_per_gpu': args.clips, 'frames_per_clip': args.frames, 'n_gpus': n_gpus,
            'sequences_in_flight_per_gpu': max(1, min(args.lanes, args.clips)),
            'ref_num': REF_NUM, 'frame_range': FRAME_RANGE, 'sigma': [SIGMA_1, SIGMA_2], 'temperature': TEMPERATURE,
            'precision': ('fp16 embeddings (VOSNet under autocast, as the reference on CUDA): one tcgen05 kind::f16 pass, '
                          'products exact in the fp32 accumulator, fp32 softmax' if args.precision == 'f16' else
                          'fp32 embeddings: bf16x3 split (hi*hi + lo*hi + hi*lo) tcgen05, fp32 accumulate/softmax'),
            'l2': f'inputs larger than L2 ({0.9 if args.precision == "f16" else 1.8:.1f} GB of embeddings per step; 126 MB L2)',
            'value_scope': 'propagation stage: append + fused affinity + merge/write-back; embeddings resident in HBM',
            'e2e_scope': 'ClipSegmenter.segment: pinned host frames -> H2D -> VOSNet(cuDNN, fp16 autocast) -> propagation -> uint8 masks -> D2H'}


class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region (NVML, ~10 ms period;
    falls back to the nvidia-smi query of B200_PROFILING.md when NVML is unavailable)."""
    REASONS = {'hw_slowdown': 0x8, 'hw_thermal_slowdown': 0x40, 'sw_thermal_slowdown': 0x20, 'sw_power_cap': 0x4}

    def __init__(self, index):
        self.index, self.sm, self.power, self.mask, self.max_sm = index, [], [], 0, None
        self._stop, self._th = threading.Event(), None

    def _physical_index(self):
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        if vis:
            ids = [v for v in vis.split(',') if v.strip()]
            if self.index < len(ids) and ids[self.index].strip().isdigit():
                return int(ids[self.index])
        return self.index

    def _run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            while not self._stop.is_set():
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                self.power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
                self.mask |= int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                self._stop.wait(0.01)
        except Exception:  # noqa: BLE001
            fields = 'clocks.sm,clocks.max.sm,power.draw'
            while not self._stop.is_set():
                try:
                    out = subprocess.run(['nvidia-smi', f'--id={self.index}', f'--query-gpu={fields}',
                                          '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                    a, b, c = [float(x) for x in out.strip().split(',')]
                    self.sm.append(a); self.max_sm = b; self.power.append(c)
                except Exception:  # noqa: BLE001
                    pass
                self._stop.wait(0.1)

    def __enter__(self):
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._th.join(timeout=6)

    def summary(self):
        if not self.sm:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_sm, 'reasons': ['unsampled']}
        return {'sm_mhz': statistics.median(self.sm), 'sm_min_mhz': min(self.sm), 'sm_max_mhz': self.max_sm,
                'power_w_median': statistics.median(self.power) if self.power else None,
                'reasons': [n for n, bit in self.REASONS.items() if self.mask & bit], 'samples': len(self.sm)}


def measured_peaks():
    p = REPO / 'MEASURED_PEAKS.json'
    if p.is_file():
        d = json.loads(p.read_text())
        return d.get('bf16_tflops_sustained', d.get('bf16_tflops')), d.get('hbm_gbs'), 'measured (MEASURED_PEAKS.json, sustained bf16)'
    return 1590.0, 6650.0, 'fallback (B200_PROFILING.md)'


def steady_refs(frame_idx):
    return min(frame_idx, REF_NUM)


# ----------------------------------------------------------------------------------------------
# CPU port of the reference (oracle) -- cpu_baseline leg and --impl reference
# ----------------------------------------------------------------------------------------------
def cpu_reference_sample(n_frames, steps, warmup):
    """End-to-end CPU frames/s of the reference algorithm on a bounded sample: `n_frames`
    steady-state frames (frame_idx >= 16, 9 references, both sigma branches) of one 480p clip per
    step; each frame = VOSNet.forward on CPU + predict + argmax/upsample.  fp32, all host threads."""
    from oracle import propagation_oracle as O
    from src.model.vos_net import VOSNet
    # all host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which would make this a 1-thread run)
    try:
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except (AttributeError, RuntimeError):
        torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    net = VOSNet('resnet50', pretrained=False).eval()
    t0_idx = 16
    T = t0_idx + n_frames
    feats, first = O.synthetic_sequence(T, H, W, 2, seed=5, feat_scale=0.30)
    low, d = O.first_frame_labels(first)
    H_d, W_d = feats.shape[2:]
    P = H_d * W_d
    g = torch.Generator().manual_seed(1)
    labels = torch.stack([O.index_to_onehot(torch.randint(0, d, (P,), generator=g), d) for _ in range(T)], 1)
    frame = torch.randn(1, 3, H, W, generator=g)
    # the reference builds the two (P,P) priors once per video (predict.py:117-118): outside the per-frame timing
    priors = (O.spatial_weight((H_d, W_d), SIGMA_1), O.spatial_weight((H_d, W_d), SIGMA_2))
    times = []
    with torch.no_grad():
        for s in range(warmup + steps):
            t0 = time.perf_counter()
            for t in range(t0_idx, T):
                _ = net(frame)                                               # P0
                pred = O.predict(feats[:t], feats[t], labels[:, :t], SIGMA_1, SIGMA_2, t, FRAME_RANGE, REF_NUM,
                                 TEMPERATURE, False, weights=priors)         # P1-P3
                up = torch.nn.functional.interpolate(pred.view(1, d, H_d, W_d), size=(H, W), mode='nearest')
                _ = torch.argmax(up, 1)                                      # P6
            if s >= warmup:
                times.append(time.perf_counter() - t0)
    per_step = sum(times) / len(times)
    return n_frames / per_step, per_step, torch.get_num_threads()


def run_reference(args, rank):
    if rank != 0:
        return
    fps, per_step, cores = cpu_reference_sample(args.ref_frames, max(args.steps, 1), min(args.warmup, 1))
    cfg = workload_config(args, args.gpus)
    sample = (f'{args.ref_frames} steady-state propagated frames (frame_idx>=16, 9 refs) of one 480p clip per step: '
              f'VOSNet.forward + predict + upsample/argmax on CPU, fp32, torch threads={cores}')
    line = {'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': per_step * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': cfg,
            'cpu_baseline': {'value': fps, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': fps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# ours
# ----------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from vosb200 import PropagationEngine
    from vosb200 import synthetic
    from vosb200.pipeline import ClipSegmenter
    from vosb200.sequence import propagate_clip, propagate_clips_lanes
    from src.model.vos_net import VOSNet

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    C, T = args.clips, args.frames
    n_obj = [2 + (i + rank) % 3 for i in range(C)]
    clips = [synthetic.clip_features(T, H, W, n_obj[i], seed=1000 * rank + i, device=dev) for i in range(C)]
    if args.precision == 'f16':
        clips = [(f.half(), first) for f, first in clips]
    passes = 1 if args.precision == 'f16' else 3
    P = clips[0][0].shape[2] * clips[0][0].shape[3]
    lanes = max(1, min(args.lanes, C))
    engines = [PropagationEngine(max_pixels=P, ring_slots=48, device=dev) for _ in range(lanes)]
    lane_streams = [torch.cuda.Stream(dev) for _ in range(lanes)]
    eng = engines[0]
    masks_keep = [None] * C

    def prop_step():
        if lanes == 1:
            for i, (feats, first) in enumerate(clips):
                masks_keep[i] = propagate_clip(eng, feats, first, SIGMA_1, SIGMA_2, FRAME_RANGE, REF_NUM, TEMPERATURE,
                                               False, d=n_obj[i] + 1)
        else:   # several sequences in flight: small kernels of one lane run under the affinity kernel of another
            masks_keep[:] = propagate_clips_lanes(engines, [(f, first, n_obj[i] + 1) for i, (f, first) in enumerate(clips)],
                                                  SIGMA_1, SIGMA_2, FRAME_RANGE, REF_NUM, TEMPERATURE, False,
                                                  streams=lane_streams)

    for _ in range(args.warmup):
        prop_step()
    launches_per_step = 3 * (T - 1) + 3        # per clip: reset + append(0) + labels(0) + (append, affinity, merge) per frame
    # inside the timed region only the dominant kernel (fused affinity) is bracketed with events: event records cost
    # front-end time (all three classes: -6 % frames/s); append / merge are timed in one extra pass afterwards
    for e_ in engines:
        e_.enable_timing(0 if args.no_kernel_events else args.steps * C * launches_per_step + 16, classes=('affinity',))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = sum(e_.launch_count for e_ in engines)
    barrier()
    with ClockSampler(local_rank) as clocks:
        ev0.record()
        for _ in range(args.steps):
            prop_step()
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    gpu_launches = sum(e_.launch_count for e_ in engines) - launches0
    stage = {'append': (0.0, 0), 'affinity': (0.0, 0), 'merge': (0.0, 0)}
    for e_ in engines:
        for k_, (ms_, n_) in e_.read_timing().items():
            stage[k_] = (stage[k_][0] + ms_, stage[k_][1] + n_)
        e_.enable_timing(0)
    if not args.no_kernel_events:
        eng.enable_timing(launches_per_step + 16, classes=('append', 'merge'))
        propagate_clip(eng, clips[0][0], clips[0][1], SIGMA_1, SIGMA_2, FRAME_RANGE, REF_NUM, TEMPERATURE, False, d=n_obj[0] + 1)
        torch.cuda.synchronize(dev)
        for k_, (ms_, n_) in eng.read_timing().items():
            stage[k_] = (stage[k_][0] + ms_, stage[k_][1] + n_)
        eng.enable_timing(0)
    t_ms = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_step = float(t_ms.item()) / args.steps
    frames_per_step = C * (T - 1) * world
    value = frames_per_step / (ms_step / 1e3)

    # roofline of the dominant kernel (fused affinity): algorithmic FLOPs 2*P*(R*P)*K per launch
    flops = sum(2.0 * P * (steady_refs(t) * P) * K for t in range(1, T)) * C * args.steps
    aff_ms, aff_n = stage['affinity']
    peak_tf, peak_hbm, peak_src = measured_peaks()
    achieved_tf = flops / (aff_ms * 1e-3) / 1e12 if aff_ms > 0 else 0.0
    roofline = {'bound': 'tensor', 'kernel': 'vos_affinity_idx', 'achieved': achieved_tf, 'peak': peak_tf, 'unit': 'TFLOP/s',
                'frac': achieved_tf / peak_tf, 'traffic': TRAFFIC_BYTES.get(args.precision), 'peak_source': peak_src,
                'issued_tflops': passes * achieved_tf, 'issued_frac': passes * achieved_tf / peak_tf,
                'launches': aff_n, 'avg_launch_us': aff_ms * 1e3 / max(aff_n, 1),
                'note': 'achieved counts algorithmic FLOPs 2*P*N*K per launch; tensor-core passes issued per logit: '
                        f'{passes}; traffic = dram bytes read+written per launch from the ncu capture in profiles/ (null: not captured)'}
    # HBM-bound side kernels: algorithmic bytes per launch
    app_ms, app_n = stage['append']
    mrg_ms, mrg_n = stage['merge']
    app_bytes = P * K * 2 * 2 if args.precision == 'f16' else P * K * 4 + 2 * P * K * 2
    mrg_bytes = 2 * 2 * 16 * 4 * P + 14 * 4 * P + H * W
    side = {'append': {'avg_launch_us': app_ms * 1e3 / max(app_n, 1), 'achieved_gbs': app_bytes * app_n / (app_ms * 1e-3) / 1e9 if app_ms else None,
                       'bytes_per_launch': app_bytes},
            'merge_writeback': {'avg_launch_us': mrg_ms * 1e3 / max(mrg_n, 1), 'achieved_gbs': mrg_bytes * mrg_n / (mrg_ms * 1e-3) / 1e9 if mrg_ms else None,
                                'bytes_per_launch': mrg_bytes},
            'hbm_peak_gbs': peak_hbm}

    # ---------------- e2e through the public API
    e2e = None
    if not args.no_e2e:
        torch.manual_seed(0)
        net = VOSNet('resnet50', pretrained=False)
        seg = ClipSegmenter(net, device=dev, sigma_1=SIGMA_1, sigma_2=SIGMA_2, frame_range=FRAME_RANGE, ref_num=REF_NUM,
                            temperature=TEMPERATURE)
        host_clips = [synthetic.clip_frames(T, H, W, n_obj[i], seed=2000 * rank + i, device=dev) for i in range(C)]
        outs = [torch.empty((T - 1, H, W), dtype=torch.uint8, pin_memory=True) for _ in range(C)]

        def e2e_step():
            for i, (frames, first) in enumerate(host_clips):
                seg.segment(frames, first, out=outs[i], sync=False)
            torch.cuda.synchronize(dev)     # results of the step are on the host
            if world > 1:                   # final per-sequence result gather (NCCL), as the north star prescribes
                res = torch.stack([o.to(dev, non_blocking=True) for o in outs])
                gathered = [torch.empty_like(res) for _ in range(world)] if rank == 0 else None
                dist.gather(res, gathered, dst=0)

        for _ in range(max(1, min(args.warmup, 2))):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {'value': frames_per_step / (float(dt.item()) / args.steps), 'unit': UNIT,
               'h2d_bytes_per_step': C * T * 3 * H * W * 4, 'd2h_bytes_per_step': C * (T - 1) * H * W,
               'ms_per_step': float(dt.item()) / args.steps * 1e3}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fps, per_step, cores = cpu_reference_sample(2, 1, 1)
        cpu = {'value': fps, 'unit': UNIT, 'cores': cores, 'kind': 'port',
               'sample': '2 steady-state propagated 480p frames (frame_idx>=16, 9 refs): VOSNet.forward + oracle predict '
                         '+ upsample/argmax on the host CPU, fp32, after 1 warm-up pass'}

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'f16' if args.precision == 'f16' else 'bf16x3', 'data': 'synthetic', 'config': workload_config(args, world),
                'clocks': clocks.summary(), 'e2e': e2e, 'gpu_launches': int(gpu_launches),
                'roofline': roofline, 'side_kernels': side, 'cpu_baseline': cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a GPU for --impl ours (there is no CPU fallback); use --impl reference on CPU')
    run_ours(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
