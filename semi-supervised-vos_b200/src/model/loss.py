"""The validation criterion on the B200 engine: same class name, constructor and call signature as the reference's
src/model/loss.py:39-66, the math on libvosprop.so instead of a (B, 9*P, P) softmax in HBM.

CrossEntropy is label propagation without a spatial prior and without frame sampling: every pixel of the first
T-1 frames of a clip votes for the last frame with weight softmax(ref . target * temperature), the loss is the NLL
of the true class under the propagated distribution (+1e-14 inside the log).  That is one `vosprop_propagate` call
per clip with all sigmas 0 -- the index-label kernel with up to 24 classes (d = 22 annotation centroids).

Only the forward value exists here (validation); training needs autograd through the affinity and is out of scope.
The focal / contrastive / triplet criteria (loss.py:69-240) depend on the miners and are not built."""
import torch
import torch.nn.functional as F
from torch import nn

from vosb200 import PropagationEngine, precision_for

_ENGINES = {}
LOG_EPS = 1e-14     # loss.py:60


def _engine(device, n_pixels, slots):
    key = (device.index, n_pixels, slots)
    if key not in _ENGINES:
        _ENGINES[key] = PropagationEngine(max_pixels=n_pixels, ring_slots=slots, device=device)
    return _ENGINES[key]


def propagate_clips(ref, target, ref_cls, d, temperature=1.0):
    """ref (B,R,K,H,W), target (B,K,H,W) CUDA fp32/fp16/bf16; ref_cls (B,R,H,W) integer class maps.  Returns the
    propagated class distribution of every target pixel, (B,d,H*W) fp32 (loss.py:55-59)."""
    B, R, K, H, W = ref.shape
    if not ref.is_cuda:
        raise RuntimeError('the propagation engine is CUDA-only: move the embeddings to the GPU (no CPU fallback)')
    eng = _engine(ref.device, H * W, 2 * (R + 1))
    eng.reset(H, W, H * 8, W * 8, d, precision_for(ref.dtype))
    out = torch.empty((B, d, H * W), dtype=torch.float32, device=ref.device)
    cls8 = ref_cls.to(device=ref.device, dtype=torch.uint8)
    zeros = [0.0] * R
    for b in range(B):
        base = (b % 2) * (R + 1)                 # alternate between two slot groups of the ring
        eng.append_frames(base, ref[b], cls8[b])     # the clip's labelled reference frames, one call
        eng.append(base + R, target[b])
        eng.propagate(base + R, list(range(base, base + R)), zeros, temperature, False, write_labels=False,
                      want_prediction=False, want_lowres=False, want_fullres=False, out_prediction=out[b])
    return out


class CrossEntropy(nn.Module):
    num_classes = 22     # rows of annotation_centroids.npy; step() overwrites it with centroids.shape[0]

    def __init__(self, temperature=1.0):
        super().__init__()
        self.temperature = temperature

    def forward(self, ref, target, ref_label, target_label, _=None, __=None, return_prediction=False):
        """ref (B,R,K,H,W); target (B,K,H,W); ref_label one-hot (B,R,d,H,W) as the reference passes it, or the
        (B,R,H,W) class maps themselves with `d` taken from `self.num_classes` (set by step()); target_label (B,H,W)."""
        if ref_label.dim() == 5:
            d, ref_cls = ref_label.shape[2], ref_label.argmax(2)
        else:
            d, ref_cls = int(self.num_classes), ref_label
        B, R, K, H, W = ref.shape
        prob = propagate_clips(ref, target, ref_cls, d, self.temperature)
        logp = torch.log(prob + LOG_EPS).reshape(B, d, H, W)
        loss = F.nll_loss(logp, target_label.to(logp.device).long())
        if return_prediction:
            return loss, logp.argmax(1)
        return loss
