"""Per-sequence propagation loop on the engine (the body of the reference's inference_single,
src/utils/inference_utils.py:27-83, with the feature extractor factored out)."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _capi as capi
from .engine import PropagationEngine, precision_for

SCALE = 0.125  # src/config.py:12


def lowres_dims(H: int, W: int) -> Tuple[int, int]:
    """src/model/predict.py:109-110"""
    return int(np.ceil(H * SCALE)), int(np.ceil(W * SCALE))


def nearest_index(out_size: int, in_size: int, device=None) -> torch.Tensor:
    """Source indices of torch's legacy 'nearest' interpolation (what F.interpolate(mode='nearest')
    does in get_labels, predict.py:94): floor(dst * (float)in/out), clamped."""
    scale = torch.tensor(float(in_size), dtype=torch.float32) / out_size
    idx = torch.floor(torch.arange(out_size, dtype=torch.float32) * scale).long().clamp_(max=in_size - 1)
    return idx.to(device) if device is not None else idx


def first_frame_lowres(label_full: torch.Tensor, H_d: int, W_d: int) -> torch.Tensor:
    """Class-index annotation (H,W) -> (H_d*W_d,) uint8.  get_labels (predict.py:92-96) one-hots at
    full resolution and nearest-downsamples; sampling the index map is the same thing."""
    H, W = label_full.shape
    ys = nearest_index(H_d, H, label_full.device)
    xs = nearest_index(W_d, W, label_full.device)
    return label_full[ys][:, xs].reshape(-1).to(torch.uint8)


def start_sequence(engine: PropagationEngine, first_features: torch.Tensor, first_label_full: torch.Tensor,
                   d: Optional[int] = None, precision: Optional[int] = None) -> int:
    """Frame 0: reset the engine, append its features, install the one-hot ground truth.
    The reference memory's precision follows the embeddings' dtype unless given (engine.precision_for)."""
    H, W = first_label_full.shape
    K, H_d, W_d = first_features.shape[-3:]
    if d is None:
        d = int(first_label_full.max().item()) + 1  # predict.py:113
    engine.reset(H_d, W_d, H, W, d, precision_for(first_features.dtype) if precision is None else precision)
    engine.append(0, first_features)
    engine.set_labels_index(0, first_frame_lowres(first_label_full.to(engine.device), H_d, W_d))
    return d


def propagate_clip(engine: PropagationEngine, features: torch.Tensor, first_label_full, sigma_1: float = 8.0,
                   sigma_2: float = 21.0, frame_range: int = 40, ref_num: int = 9, temperature: float = 1.0,
                   probability_propagation: bool = False, kernel: int = capi.KERNEL_TC, d: Optional[int] = None,
                   return_predictions: bool = False, precision: Optional[int] = None, topk: int = 0):
    """features (T,K,H_d,W_d) on the engine's device -> masks (T-1,H,W) uint8 on device
    (+ predictions (T-1,d,P) fp32 when asked).  No host sync inside."""
    first = torch.as_tensor(np.asarray(first_label_full)) if not torch.is_tensor(first_label_full) else first_label_full
    H, W = first.shape
    T = features.shape[0]
    d = start_sequence(engine, features[0], first, d, precision)
    P = features.shape[2] * features.shape[3]
    masks = torch.empty((max(T - 1, 0), H, W), dtype=torch.uint8, device=engine.device)
    preds = torch.empty((T - 1, d, P), dtype=torch.float32, device=engine.device) if return_predictions else None
    ahead = engine.lookahead(frame_range, ref_num)
    appended = 1                         # frames 0 .. appended-1 are in the ring
    for t in range(1, T):
        if t >= appended:                # one launch appends the next `ahead` frames: off the per-frame chain
            n = min(ahead, T - appended)
            engine.append_frames(appended, features[appended:appended + n])
            appended += n
        engine.step(t, frame_range, ref_num, sigma_1, sigma_2, temperature, probability_propagation,
                    kernel=kernel, want_prediction=False, want_lowres=False, want_fullres=False,
                    out_fullres=masks[t - 1], out_prediction=preds[t - 1] if return_predictions else None, topk=topk)
    return (masks, preds) if return_predictions else masks


def propagate_clips_lanes(engines: Sequence[PropagationEngine], clips, sigma_1: float = 8.0, sigma_2: float = 21.0,
                          frame_range: int = 40, ref_num: int = 9, temperature: float = 1.0,
                          probability_propagation: bool = False, kernel: int = capi.KERNEL_TC, topk: int = 0,
                          streams: Optional[Sequence[torch.cuda.Stream]] = None):
    """Several sequences in flight on one GPU: clip i runs on lane i % len(engines) (one engine + one stream per lane).

    Frames inside a sequence are strictly serial (the labels of frame t feed frame t+1), and the fused affinity kernel
    of one sequence fills every SM.  The lanes advance round-robin, one frame each, and their affinity kernels are
    chained through events (vosprop_step.wait_event / record_event) so they run back to back on the device while the
    other lanes' launch latencies and small kernels fill the gaps.  (Measured in round 2, DESIGN.md section 10: more sequences in
    flight do not pay on this GPU -- merges that run under the other sequence's fused kernel are won back by its slower tiles,
    two reference memories no longer fit the part of the L2 one die can use.)  clips: [(features (T,K,H_d,W_d), first annotation (H,W), d or None), ...].
    Returns one (T-1,H,W) uint8 device tensor per clip.  No host sync inside; the current stream waits for the lanes."""
    n_lanes = len(engines)
    dev = engines[0].device
    main = torch.cuda.current_stream(dev)
    streams = list(streams) if streams is not None else [torch.cuda.Stream(dev) for _ in range(n_lanes)]
    events = [torch.cuda.Event() for _ in range(n_lanes)]
    for s, ev in zip(streams, events):
        s.wait_stream(main)
        ev.record(s)                      # creates the CUDA event behind torch's lazy wrapper
    lanes = [[i for i in range(len(clips)) if i % n_lanes == lane] for lane in range(n_lanes)]
    outs = [None] * len(clips)
    cursor = [[0, 0] for _ in range(n_lanes)]         # per lane: (position in its clip list, next frame)
    appended = [1] * n_lanes                          # per lane: frames of its current clip already in the ring
    prev_event = None
    busy = True
    while busy:
        busy = False
        for lane in range(n_lanes):
            pos, t = cursor[lane]
            if pos >= len(lanes[lane]):
                continue
            busy = True
            ci = lanes[lane][pos]
            feats, first, d = clips[ci]
            eng = engines[lane]
            with torch.cuda.stream(streams[lane]):
                if t == 0:
                    first_t = torch.as_tensor(np.asarray(first)) if not torch.is_tensor(first) else first
                    start_sequence(eng, feats[0], first_t, d)
                    H, W = first_t.shape
                    outs[ci] = torch.empty((feats.shape[0] - 1, H, W), dtype=torch.uint8, device=dev)
                    appended[lane] = 1
                else:
                    if t >= appended[lane]:           # one launch appends the next frames (engine.lookahead)
                        n = min(eng.lookahead(frame_range, ref_num), feats.shape[0] - appended[lane])
                        eng.append_frames(appended[lane], feats[appended[lane]:appended[lane] + n])
                        appended[lane] += n
                    eng.step(t, frame_range, ref_num, sigma_1, sigma_2, temperature, probability_propagation, kernel=kernel,
                             want_prediction=False, want_lowres=False, want_fullres=False, out_fullres=outs[ci][t - 1],
                             topk=topk, wait_event=prev_event, record_event=events[lane])
                    prev_event = events[lane]
            t += 1
            cursor[lane] = [pos + 1, 0] if t == feats.shape[0] else [pos, t]
    for s in streams:
        main.wait_stream(s)
    return outs

