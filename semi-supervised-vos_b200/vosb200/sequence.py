"""Per-sequence propagation loop on the engine (the body of the reference's inference_single,
src/utils/inference_utils.py:27-83, with the feature extractor factored out)."""
from __future__ import annotations

from typing import Optional, Tuple

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
    for t in range(1, T):
        engine.append(t, features[t])
        engine.step(t, frame_range, ref_num, sigma_1, sigma_2, temperature, probability_propagation,
                    kernel=kernel, want_prediction=False, want_lowres=False, want_fullres=False,
                    out_fullres=masks[t - 1], out_prediction=preds[t - 1] if return_predictions else None, topk=topk)
    return (masks, preds) if return_predictions else masks
