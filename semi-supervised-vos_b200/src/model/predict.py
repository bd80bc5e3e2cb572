"""Drop-in for the reference's src/model/predict.py: same callables, same signatures, the math on
the B200 engine (libvosprop.so) instead of torch ops.

* `predict(...)`  -- the 10-argument, stateless signature of predict.py:19-28.  The caller passes the
  whole history; this adapter stages the sampled reference frames into a scratch engine and runs
  one fused propagation step.  (The streaming loop in src/utils/inference_utils.py keeps an engine
  per sequence instead and never re-stages history.)
* `get_spatial_weight` returns a `SpatialPrior` descriptor -- (sigma, H_d, W_d) -- instead of the
  (P,P) matrix (165 MB at 480p, 4.2 GB at 1080p): the kernel evaluates the Gaussian in closed form.
  `SpatialPrior.materialize()` builds the explicit matrix for comparisons only.
"""
import os
from typing import Optional

import numpy as np
import torch
from PIL import Image

from src.config import Config
from vosb200 import PropagationEngine, plan_refs, precision_for
from vosb200 import sample_frames as _sample_frames
from vosb200.sequence import first_frame_lowres

_SCRATCH = {}


class SpatialPrior:
    """Descriptor of W[i,j] = exp(-((i/W_d - j/W_d)^2 + (i%W_d - j%W_d)^2) / sigma^2) (predict.py:158-175)."""

    def __init__(self, shape, sigma):
        self.shape = (int(shape[0]), int(shape[1]))
        self.sigma = float(sigma)

    def materialize(self, device=None):
        H, W = self.shape
        idx = torch.arange(H * W, dtype=torch.long, device=device)
        row, col = idx.div(float(W)), (idx % W).float()
        d2 = (row[:, None] - row[None, :]).pow(2) + (col[:, None] - col[None, :]).pow(2)
        return (-d2 / self.sigma ** 2).exp()

    def __repr__(self):
        return f'SpatialPrior(shape={self.shape}, sigma={self.sigma})'


def _sigma_of(weight, W_d) -> float:
    """sigma from a descriptor, or recovered from an explicit (P,P) matrix via W[0,1] = exp(-(1/W_d^2 + 1)/sigma^2)."""
    if weight is None:
        return 0.0
    if isinstance(weight, SpatialPrior):
        return weight.sigma
    w01 = float(weight[0, 1])
    if not 0.0 < w01 < 1.0:
        raise ValueError('cannot recover sigma from the given spatial weight matrix')
    return float(np.sqrt(-(1.0 + 1.0 / (W_d * W_d)) / np.log(w01)))


def sample_frames(frame_idx, take_range, num_refs):
    """predict.py:74-89 -> LongTensor on Config.DEVICE (host arithmetic lives in the C library)."""
    return torch.tensor(_sample_frames(frame_idx, take_range, num_refs), dtype=torch.long, device=Config.DEVICE)


def predict(ref, target, ref_label, weight_dense, weight_sparse, frame_idx, range, ref_num, temperature,
            probability_propagation):
    """ref (T,K,H_d,W_d), target (K,H_d,W_d), ref_label (d,T,P) -> prediction (d,P) fp32."""
    d = ref_label.shape[0]
    K, H_d, W_d = target.shape
    P = H_d * W_d
    frames, _ = plan_refs(frame_idx, range, ref_num, 1.0, 1.0, probability_propagation)
    R = len(frames)
    dev = target.device
    key = (dev.index, P)
    eng = _SCRATCH.get(key)
    if eng is None:
        eng = _SCRATCH[key] = PropagationEngine(max_pixels=P, ring_slots=33, device=dev)
    if ref.dtype != target.dtype:
        ref = ref.to(target.dtype)
    eng.reset(H_d, W_d, H_d * 8, W_d * 8, d, precision_for(target.dtype))
    for slot, f in enumerate(frames):
        eng.append(slot, ref[f])
        eng.set_labels_dense(slot, ref_label[:, f])
    eng.append(R, target)
    if probability_propagation:
        sigmas = [0.0] * R
    else:
        s_dense, s_sparse = _sigma_of(weight_dense, W_d), _sigma_of(weight_sparse, W_d)
        n_sparse = max(R - Config.CONTINUOUS_FRAME, 0) if frame_idx > 15 else 0   # predict.py:60-66
        sigmas = [s_sparse] * n_sparse + [s_dense] * (R - n_sparse)
    out = eng.propagate(R, list(np.arange(R)), sigmas, temperature, probability_propagation, write_labels=False,
                        want_lowres=False, want_fullres=False)
    return out['prediction']


def get_labels(label, d, H, W, H_d, W_d):
    """(H,W) class indices -> (d,1,H_d*W_d) int32 one-hot, nearest down-sampled (predict.py:92-96)."""
    low = first_frame_lowres(label.view(H, W), H_d, W_d).long()
    one = torch.zeros(d, H_d * W_d, dtype=torch.int32, device=low.device).scatter_(0, low.view(1, -1), 1)
    return one.unsqueeze(1)


def get_spatial_weight(shape, sigma, t_loc: Optional[float] = None):
    if t_loc is not None:
        raise NotImplementedError('t_loc is never used by the reference (predict.py:170-171)')
    return SpatialPrior(shape, sigma)


def prepare_first_frame(curr_video, save_prediction, annotation, sigma1=8, sigma2=21, inference_strategy='single',
                        probability_propagation=False, scale=None):
    """Read the first annotation (palette PNG), derive d / palette / low-res one-hot labels and the
    prior descriptors, and copy the annotation to `<save>/<video>/00000.png` (predict.py:99-155)."""
    first_annotation = Image.open(annotation)
    label_np = np.asarray(first_annotation)
    H, W = label_np.shape
    H_d, W_d = int(np.ceil(H * Config.SCALE)), int(np.ceil(W * Config.SCALE))
    palette = first_annotation.getpalette()
    d = int(label_np.max()) + 1
    label = torch.from_numpy(label_np.astype(np.int64)).to(Config.DEVICE)
    label_1hot = get_labels(label, d, H, W, H_d, W_d)
    prior = (lambda shp, s: None) if probability_propagation else get_spatial_weight
    weight_dense, weight_sparse = prior((H_d, W_d), sigma1), prior((H_d, W_d), sigma2)
    if save_prediction is not None:
        save_path = os.path.join(save_prediction, curr_video)
        os.makedirs(save_path, exist_ok=True)
        first_annotation.save(os.path.join(save_path, '00000.png'))
    if inference_strategy in ('hor-flip', 'ver-flip'):
        flip = torch.fliplr if inference_strategy == 'hor-flip' else torch.flipud
        return label_1hot, get_labels(flip(label), d, H, W, H_d, W_d), d, palette, weight_dense, weight_sparse
    if inference_strategy in ('2-scale', 'hor-2-scale', '3-scale'):
        H_2, W_2 = int(np.ceil(H * Config.SCALE * scale)), int(np.ceil(W * Config.SCALE * scale))
        dense_2, sparse_2 = prior((H_2, W_2), sigma1), prior((H_2, W_2), sigma2)
        label_2 = get_labels(label, d, H, W, H_2, W_2)
        if inference_strategy == '3-scale':
            return label_2, d, palette, dense_2, sparse_2
        return (label_1hot, label_2), d, palette, (weight_dense, dense_2), (weight_sparse, sparse_2)
    return label_1hot, d, palette, weight_dense, weight_sparse
