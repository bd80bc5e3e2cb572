"""CPU oracle for the label-propagation hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``semi-supervised-vos_b200/``) never does: it calls the CUDA extension through the C ABI and
fails loudly when the extension is missing.

This is a restatement, in plain torch-CPU fp32 / numpy, of the reference's algorithm.  Every
function cites the reference file:line it follows (paths relative to the reference root).
Parity pin: the reference ships no tests or golden vectors, so the pins are
produced by *executing the reference's own functions* in the build container
(``oracle/make_golden.py`` -> ``tests/golden/*.npz``) and ``tests/test_oracle_golden.py``
checks this restatement against them.

Where the reference cannot run (1080p affinity = 37.8 GB, top-k which the reference lacks) the
restatement is column-chunked / extended; the chunked form is validated against the
un-chunked one at sizes the reference does run.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

CONTINUOUS_FRAME = 4  # src/config.py:13
SCALE = 0.125  # src/config.py:12
DENSE_SWITCH_FRAME = 15  # src/model/predict.py:60  (`frame_idx > 15`)


# --------------------------------------------------------------------------------------
# P1  sample_frames  (src/model/predict.py:74-89)
# --------------------------------------------------------------------------------------
def sample_frames(frame_idx: int, take_range: int, num_refs: int) -> List[int]:
    """History indices used as references for target frame ``frame_idx``.

    src/model/predict.py:77-87.  ``np.linspace(...).astype(int)`` truncates toward zero; the
    reference raises ValueError for num_refs < 3 once frame_idx > num_refs (negative
    ``sparse_num``) and we keep that behaviour.
    """
    if frame_idx <= num_refs:
        return list(range(frame_idx))
    dense_num = CONTINUOUS_FRAME - 1
    sparse_num = num_refs - dense_num
    ref_end = frame_idx - dense_num - 1
    ref_start = max(ref_end - take_range, 0)
    idx = np.linspace(ref_start, ref_end, sparse_num).astype(int).tolist()
    idx.extend(frame_idx - dense_num + j for j in range(dense_num))
    return idx


def max_lookback(take_range: int) -> int:
    """Furthest history distance sample_frames can reach: dense_num + 1 + take_range."""
    return CONTINUOUS_FRAME + take_range


# --------------------------------------------------------------------------------------
# P2  get_spatial_weight  (src/model/predict.py:158-175)
# --------------------------------------------------------------------------------------
def pixel_coords(H: int, W: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(row, col) coordinates exactly as the reference builds them.

    src/model/predict.py:167-168: ``index.div(float(W))`` is TRUE division of a LongTensor by a
    python float -> fp32 fractional row; ``index % W`` is the integer column (promoted to fp32
    by the cat).
    """
    idx = torch.arange(H * W, dtype=torch.long)
    return idx.div(float(W)), (idx % W).float()


def spatial_weight(shape: Tuple[int, int], sigma: float,
                   cols: Optional[slice] = None) -> torch.Tensor:
    """W[i, j] = exp(-((row_i-row_j)^2 + (x_i-x_j)^2) / sigma^2), optionally only columns ``cols``.

    Same fp32 operation order as src/model/predict.py:169-173 (sub, pow(2), sum, neg, div, exp)
    without the (P,P,2) intermediate.  W is symmetric, so [ref_pixel, target_pixel] indexing in
    predict() equals [target, ref].
    """
    H, W = shape
    row, col = pixel_coords(H, W)
    rj, cj = (row, col) if cols is None else (row[cols], col[cols])
    # reference: d = index_matrix - index_matrix.unsqueeze(1): d[a, b] = coord[b] - coord[a]
    # restricted to columns: out[i, j] with i over all pixels, j over `cols`.
    dr = row.unsqueeze(1) - rj.unsqueeze(0)
    dc = col.unsqueeze(1) - cj.unsqueeze(0)
    d = dr.pow(2) + dc.pow(2)
    return (-d / sigma ** 2).exp()


# --------------------------------------------------------------------------------------
# P4  index_to_onehot (src/utils/utils.py:59-68)  /  get_labels (src/model/predict.py:92-96)
# --------------------------------------------------------------------------------------
def index_to_onehot(idx: torch.Tensor, d: int) -> torch.Tensor:
    n = idx.shape[0]
    return torch.zeros(d, n).scatter_(0, idx.view(1, -1).long(), 1)


def lowres_dims(H: int, W: int) -> Tuple[int, int]:
    """src/model/predict.py:109-110."""
    return int(np.ceil(H * SCALE)), int(np.ceil(W * SCALE))


def nearest_src_index(out_size: int, in_size: int) -> np.ndarray:
    """Source index of torch's legacy 'nearest' interpolation: floor(dst * in/out) in fp32.

    (ATen nearest_neighbor_compute_source_index: scale = (float)in/out; min(floor(dst*scale), in-1).)
    """
    scale = np.float32(in_size) / np.float32(out_size)
    src = np.floor(np.arange(out_size, dtype=np.float32) * scale).astype(np.int64)
    return np.minimum(src, in_size - 1)


def first_frame_labels(label_full: np.ndarray, d: Optional[int] = None,
                       dims: Optional[Tuple[int, int]] = None) -> Tuple[torch.Tensor, int]:
    """Full-res class-index annotation (H,W) -> low-res class index (P,) and d.

    src/model/predict.py:107-114 + get_labels :92-96 (one-hot at full res, nearest down-sample
    to (H_d, W_d), int32).  Nearest sampling of a one-hot tensor equals nearest sampling of the
    index map, which is what we return (the one-hot is index_to_onehot of it).
    """
    H, W = label_full.shape
    if d is None:
        d = int(label_full.max()) + 1  # predict.py:113
    H_d, W_d = lowres_dims(H, W) if dims is None else dims      # dims: second scale of the 2-scale strategies (predict.py:137-141)
    ys = nearest_src_index(H_d, H)
    xs = nearest_src_index(W_d, W)
    low = torch.from_numpy(np.ascontiguousarray(label_full[np.ix_(ys, xs)]).astype(np.int64))
    return low.reshape(-1), d


def upsample_mask(low_mask: torch.Tensor, H_d: int, W_d: int, H: int, W: int) -> torch.Tensor:
    """argmax-at-stride-8 then nearest up-sample == the reference's up-sample then argmax
    (src/utils/inference_utils.py:74-75): nearest replication commutes with argmax."""
    ys = torch.from_numpy(nearest_src_index(H, H_d))
    xs = torch.from_numpy(nearest_src_index(W, W_d))
    return low_mask.view(H_d, W_d)[ys][:, xs]


# --------------------------------------------------------------------------------------
# P3  predict  (src/model/predict.py:19-71)
# --------------------------------------------------------------------------------------
def predict(ref: torch.Tensor, target: torch.Tensor, ref_label: torch.Tensor,
            sigma_dense: Optional[float], sigma_sparse: Optional[float], frame_idx: int,
            take_range: int, ref_num: int, temperature: float,
            probability_propagation: bool, chunk: Optional[int] = None,
            topk: Optional[int] = None, return_topk_idx: bool = False,
            weights: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
            pixel_range: Optional[Tuple[int, int]] = None, cuda_half: bool = False):
    """Label propagation for one target frame.

    ref (T,K,H,W) fp32, target (K,H,W), ref_label (d,T,P) -> prediction (d,P) fp32.
    Follows src/model/predict.py:40-70 line by line; the (P,P) priors are rebuilt per column
    chunk from (sigma, H, W) instead of being passed in.

    ``weights``: optional precomputed (weight_dense, weight_sparse) (P,P) matrices, as the
    reference passes them (built once per video by prepare_first_frame, predict.py:117-118).
    ``chunk``: process target pixels in column blocks of this size (softmax over dim 0 is
    per-column, so chunking is exact up to GEMM blocking).
    ``pixel_range``: (p0, p1) -- evaluate only target pixels p0..p1-1 and return (d, p1-p0); every target
    pixel is independent (softmax over dim 0), so this is the corresponding column block of the full result.
    Used where the full product is too large for a CPU test (1080p: 291 600 x 32 400).
    ``topk``: EXTENSION (not in the reference): the softmax of predict.py:55 is
    restricted, per target pixel, to the k reference pixels with the largest logit (ties ->
    lowest reference index first); everything else gets weight 0.  The prior and label gather
    are unchanged, so topk >= N is exactly the reference.
    ``cuda_half``: emulate the rounding of the reference's own CUDA path, where the embeddings leave the network in fp16
    (autocast, inference_utils.py:52-53) and predict() runs OUTSIDE autocast on them: `ref.mm(target)` returns fp16 logits
    (fp32 accumulate, one rounding), `*= temperature` and the softmax are fp16 (computed in fp32, rounded on output), the
    in-place `*=` with the prior keeps fp16 (frame_idx > 15) while the out-of-place `.mul` promotes to fp32 (else branch),
    and `.float()` precedes the label product (predict.py:49-70).  The product path of this repository keeps fp32 logits and
    softmax on the same fp16 inputs (the more accurate of the two); tests/test_oracle_golden.py reports how far apart the
    masks of the two are.
    """
    d = ref_label.shape[0]
    idx = torch.tensor(sample_frames(frame_idx, take_range, ref_num), dtype=torch.long)
    ref_sel = ref.index_select(0, idx)                                   # :42
    lab_sel = ref_label.index_select(1, idx).reshape(d, -1).float()      # :43, :70
    R, K, H, W = ref_sel.shape
    P = H * W
    ref_mat = ref_sel.permute(0, 2, 3, 1).reshape(-1, K)                 # :47  (N,K)
    tgt = target.reshape(K, -1)                                          # :48
    out = torch.empty(d, P, dtype=torch.float32)
    topk_idx = torch.empty(P, topk, dtype=torch.long) if (topk and return_topk_idx) else None
    step = P if chunk is None else chunk
    use_prior = not probability_propagation
    p0, p1 = (0, P) if pixel_range is None else pixel_range
    for c0 in range(p0, p1, step):
        cs = slice(c0, min(c0 + step, p1))
        S = ref_mat.mm(tgt[:, cs])                                       # :49
        if cuda_half:
            S = S.half()
            S *= temperature
            S = S.float().softmax(dim=0).half()
            if use_prior:
                S = S.view(R, P, -1)
                w_dense = weights[0][:, cs] if weights else spatial_weight((H, W), sigma_dense, cs)
                if frame_idx > DENSE_SWITCH_FRAME:
                    w_sparse = weights[1][:, cs] if weights else spatial_weight((H, W), sigma_sparse, cs)
                    S[:-CONTINUOUS_FRAME] = (S[:-CONTINUOUS_FRAME].float() * w_sparse).half()
                    S[-CONTINUOUS_FRAME:] = (S[-CONTINUOUS_FRAME:].float() * w_dense).half()
                else:
                    S = S.float().mul(w_dense)
                S = S.reshape(R * P, -1)
            out[:, cs] = lab_sel.mm(S.float())
            continue
        S *= temperature                                                 # :52
        if topk is not None and topk < S.shape[0]:
            # stable descending sort => ties resolved toward the lowest reference index
            order = torch.sort(S, dim=0, descending=True, stable=True).indices[:topk]
            kept = torch.gather(S, 0, order).softmax(dim=0)
            S = torch.zeros_like(S).scatter_(0, order, kept)
            if topk_idx is not None:
                topk_idx[cs] = order.t()
        else:
            S = S.softmax(dim=0)                                         # :55
            if topk_idx is not None:
                topk_idx[cs] = torch.sort(S, dim=0, descending=True, stable=True).indices[:topk].t()
        if use_prior:                                                    # :59-66
            S = S.view(R, P, -1)
            w_dense = weights[0][:, cs] if weights else spatial_weight((H, W), sigma_dense, cs)
            if frame_idx > DENSE_SWITCH_FRAME:
                w_sparse = weights[1][:, cs] if weights else spatial_weight((H, W), sigma_sparse, cs)
                S[:-CONTINUOUS_FRAME] *= w_sparse
                S[-CONTINUOUS_FRAME:] *= w_dense
            else:
                S = S.mul(w_dense)
            S = S.reshape(R * P, -1)
        out[:, cs] = lab_sel.mm(S.float())                               # :70
    if pixel_range is not None:
        out = out[:, p0:p1]
        topk_idx = topk_idx[p0:p1] if topk_idx is not None else None
    if return_topk_idx:
        return out, topk_idx
    return out


def ref_sigmas(frame_idx: int, n_refs: int, sigma_dense: float, sigma_sparse: float
               ) -> List[float]:
    """Per-reference sigma implied by src/model/predict.py:60-66 (negative-slice semantics:
    with R <= 4 refs ``S[:-4]`` is empty and every ref gets the dense sigma)."""
    if frame_idx > DENSE_SWITCH_FRAME:
        n_sparse = max(n_refs - CONTINUOUS_FRAME, 0)
        return [sigma_sparse] * n_sparse + [sigma_dense] * (n_refs - n_sparse)
    return [sigma_dense] * n_refs


# --------------------------------------------------------------------------------------
# P5/P6/P8  the per-sequence loop  (src/utils/inference_utils.py:23-87)
# --------------------------------------------------------------------------------------
def propagate_sequence(features: torch.Tensor, first_label_full: np.ndarray,
                       sigma_1: float = 8.0, sigma_2: float = 21.0, frame_range: int = 40,
                       ref_num: int = 9, temperature: float = 1.0,
                       probability_propagation: bool = False, chunk: Optional[int] = None,
                       d: Optional[int] = None, topk: Optional[int] = None, cuda_half: bool = False):
    """inference_single with the feature extractor factored out.

    features (T,K,H_d,W_d) fp32 = model(frame_t) for every frame of one video.
    Returns (masks (T-1,H,W) int64, predictions list of (d,P) fp32).
    Memory is the reference's unbounded cat (inference_utils.py:71-72); the product uses a ring
    and tests/test_ring.py shows the two are equivalent.
    """
    H, W = first_label_full.shape
    T, K, H_d, W_d = features.shape
    low, d = first_frame_labels(first_label_full, d)
    label_history = index_to_onehot(low, d).unsqueeze(1)                 # (d,1,P)  predict.py:93-96
    feats_history = features[:1]
    masks, preds = [], []
    for t in range(1, T):
        pred = predict(feats_history, features[t], label_history, sigma_1, sigma_2, t,
                       frame_range, ref_num, temperature, probability_propagation, chunk, topk, cuda_half=cuda_half)
        if probability_propagation:                                      # :67-70
            new_label = pred.unsqueeze(1)
        else:
            new_label = index_to_onehot(torch.argmax(pred, 0), d).unsqueeze(1)
        label_history = torch.cat((label_history, new_label), 1)         # :71
        feats_history = torch.cat((feats_history, features[t:t + 1]), 0)  # :72
        up = torch.nn.functional.interpolate(pred.view(1, d, H_d, W_d), size=(H, W),
                                             mode='nearest')             # :74
        masks.append(torch.argmax(up, 1)[0])                             # :75
        preds.append(pred)
    return torch.stack(masks), preds


# --------------------------------------------------------------------------------------
# Test-time-augmentation strategies: two independent memories + per-frame fusion
# (src/utils/inference_utils.py:90-511)
# --------------------------------------------------------------------------------------
def _propagate_stream(features, low, d, out_hw, sigma_1, sigma_2, frame_range, ref_num, temperature,
                      probability_propagation):
    """One memory of a two-stream strategy: per frame the prediction nearest-up-sampled to out_hw, (1,d,H,W)."""
    T, K, H_d, W_d = features.shape
    label_history = index_to_onehot(low, d).unsqueeze(1)
    feats_history = features[:1]
    ups = []
    for t in range(1, T):
        pred = predict(feats_history, features[t], label_history, sigma_1, sigma_2, t, frame_range, ref_num,
                       temperature, probability_propagation)
        new_label = pred.unsqueeze(1) if probability_propagation else index_to_onehot(torch.argmax(pred, 0), d).unsqueeze(1)
        label_history = torch.cat((label_history, new_label), 1)
        feats_history = torch.cat((feats_history, features[t:t + 1]), 0)
        ups.append(torch.nn.functional.interpolate(pred.view(1, d, H_d, W_d), size=out_hw, mode='nearest'))
    return ups


def propagate_two_streams(strategy: str, feats_a: torch.Tensor, feats_b: torch.Tensor, first_label_full: np.ndarray,
                          sigma_1: float = 8.0, sigma_2: float = 21.0, frame_range: int = 40, ref_num: int = 9,
                          temperature: float = 1.0, probability_propagation: bool = False, reduction: str = 'mean',
                          scale: float = 1.15) -> torch.Tensor:
    """inference_hor_flip (:90-192), inference_ver_flip (:195-298), inference_2_scale (:302-410, incl. 'hor-2-scale'),
    inference_multimodel (:411-511) with the feature extractor factored out -> fused masks (T-1,H,W) uint8.
    Quirks kept: both flip strategies un-flip stream B with torch.fliplr (:173, :279), which on the (1,d,H,W)
    probability map flips the class axis; 'hor-2-scale' feeds stream B unflipped labels (predict.py:136-141);
    probabilities are cast to half before the arg-max (:181)."""
    H, W = first_label_full.shape
    lab = np.asarray(first_label_full)
    low_a, d = first_frame_labels(lab)
    if strategy == 'hor-flip':
        low_b, _ = first_frame_labels(np.ascontiguousarray(lab[:, ::-1]), d)
    elif strategy == 'vert-flip':
        low_b, _ = first_frame_labels(np.ascontiguousarray(lab[::-1]), d)
    elif strategy in ('2-scale', 'hor-2-scale'):
        low_b, _ = first_frame_labels(lab, d, dims=(int(np.ceil(H * SCALE * scale)), int(np.ceil(W * SCALE * scale))))
    elif strategy == 'multimodel':
        low_b = low_a
    else:
        raise ValueError(strategy)
    args = (sigma_1, sigma_2, frame_range, ref_num, temperature, probability_propagation)
    ups_a = _propagate_stream(feats_a, low_a, d, (H, W), *args)
    ups_b = _propagate_stream(feats_b, low_b, d, (H, W), *args)
    reduce = {'maximum': torch.maximum, 'minimum': torch.minimum, 'mean': lambda x, y: (x + y) / 2.0}[reduction]
    out = []
    for pa, pb in zip(ups_a, ups_b):
        if not probability_propagation:
            pa, pb = torch.argmax(pa, 1)[0], torch.argmax(pb, 1)[0]          # (H,W)
        if strategy in ('hor-flip', 'vert-flip'):
            pb = torch.fliplr(pb)
        elif strategy == 'hor-2-scale':
            pb = torch.flip(pb, dims=(-1,))
        if probability_propagation:
            out.append(torch.argmax(reduce(pa, pb).half(), 1)[0])
        else:
            out.append(torch.maximum(pa, pb))
    return torch.stack(out).to(torch.uint8)


THREE_SCALE_OUT = (480, 910)     # inference_utils.py:574 -- hard-coded, whatever the input size


def propagate_three_scales(feats_by_scale: Sequence[torch.Tensor], first_label_full: np.ndarray, scale: float = 1.15,
                           sigma_1: float = 8.0, sigma_2: float = 21.0, frame_range: int = 40, ref_num: int = 9,
                           temperature: float = 1.0, probability_propagation: bool = False) -> torch.Tensor:
    """inference_3_scale (src/utils/inference_utils.py:514-595) with the feature extractor factored out: three
    independent single-stream propagations on inputs nearest-resized by 0.9 / 1.0 / `scale` (first-frame labels sampled
    at ceil(H * 0.125 * s), predict.py:146-153), every prediction nearest-up-sampled to 480 x 910 and arg-maxed, the three
    label maps fused by an element-wise maximum of the class indices (:594).  -> (T-1, 480, 910) uint8."""
    H, W = first_label_full.shape
    lab = np.asarray(first_label_full)
    _, d = first_frame_labels(lab)
    fused = None
    for feats, s in zip(feats_by_scale, (0.9, 1.0, scale)):
        low, _ = first_frame_labels(lab, d, dims=(int(np.ceil(H * SCALE * s)), int(np.ceil(W * SCALE * s))))
        ups = _propagate_stream(feats, low, d, THREE_SCALE_OUT, sigma_1, sigma_2, frame_range, ref_num, temperature,
                                probability_propagation)
        masks = torch.stack([torch.argmax(u, 1)[0] for u in ups])
        fused = masks if fused is None else torch.maximum(fused, masks)
    return fused.to(torch.uint8)


# --------------------------------------------------------------------------------------
# bf16 hi/lo split emulation (what the tcgen05 kernel feeds the tensor cores) -- used by tests
# to bound the error budget, not a reference behaviour.
# --------------------------------------------------------------------------------------
def split_bf16(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi, lo


# --------------------------------------------------------------------------------------
# Synthetic, seeded inputs shared by tests / smoke / bench (no dataset, no network).
# --------------------------------------------------------------------------------------
def synthetic_sequence(T: int, H: int, W: int, n_objects: int, K: int = 256, seed: int = 0,
                       noise: float = 0.35, feat_scale: float = 1.0
                       ) -> Tuple[torch.Tensor, np.ndarray]:
    """A moving-blobs clip expressed directly in embedding space.

    Returns (features (T,K,H_d,W_d) fp32, first-frame annotation (H,W) uint8 class indices).
    Each object (and the background) owns a random unit-ish embedding; object k is an ellipse
    that drifts over time; a smooth texture field and white noise are added so that logits have
    a realistic spread (std ~ tens) and masks keep several live classes -- unlike random-init
    ResNet features, which are degenerate.
    """
    g = torch.Generator().manual_seed(seed)
    H_d, W_d = lowres_dims(H, W)
    proto = torch.randn(n_objects + 1, K, generator=g) * feat_scale
    tex_basis = torch.randn(8, K, generator=g) * (0.5 * feat_scale)
    yy, xx = torch.meshgrid(torch.arange(H_d, dtype=torch.float32),
                            torch.arange(W_d, dtype=torch.float32), indexing='ij')
    cy = torch.rand(n_objects, generator=g) * 0.6 + 0.2
    cx = torch.rand(n_objects, generator=g) * 0.6 + 0.2
    ry = torch.rand(n_objects, generator=g) * 0.12 + 0.10
    rx = torch.rand(n_objects, generator=g) * 0.12 + 0.10
    vy = (torch.rand(n_objects, generator=g) - 0.5) * 0.03
    vx = (torch.rand(n_objects, generator=g) - 0.5) * 0.03
    phase = torch.rand(8, 2, generator=g) * 6.28
    freq = torch.rand(8, 2, generator=g) * 0.35 + 0.05

    def class_map(t: int, hh: int, ww: int) -> torch.Tensor:
        y = (torch.arange(hh, dtype=torch.float32) + 0.5) / hh
        x = (torch.arange(ww, dtype=torch.float32) + 0.5) / ww
        Y, X = torch.meshgrid(y, x, indexing='ij')
        cm = torch.zeros(hh, ww, dtype=torch.long)
        for k in range(n_objects):
            inside = ((Y - (cy[k] + vy[k] * t)) / ry[k]) ** 2 + ((X - (cx[k] + vx[k] * t)) / rx[k]) ** 2 <= 1.0
            cm[inside] = k + 1
        return cm

    feats = torch.empty(T, K, H_d, W_d)
    for t in range(T):
        cm = class_map(t, H_d, W_d)
        base = proto[cm]                                                  # (H_d,W_d,K)
        coef = torch.stack([torch.sin(freq[i, 0] * (yy + 0.7 * t) + phase[i, 0]) *
                            torch.cos(freq[i, 1] * (xx - 0.4 * t) + phase[i, 1])
                            for i in range(8)], -1)                       # (H_d,W_d,8)
        f = base + coef @ tex_basis + noise * feat_scale * torch.randn(H_d, W_d, K, generator=g)
        feats[t] = f.permute(2, 0, 1)
    first = class_map(0, H, W).numpy().astype(np.uint8)
    return feats, first
