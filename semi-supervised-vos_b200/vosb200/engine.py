"""PropagationEngine: torch-facing wrapper of one libvosprop handle.

PyTorch is used for device memory and streams only; all arithmetic of the propagation path runs
in the CUDA kernels of libvosprop.so (csrc/).  One engine per video sequence in flight.  Every
method is stream-ordered on torch's current stream and never synchronises.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch

from . import _capi as capi

CONTINUOUS_FRAME = 4          # src/config.py:13
DEFAULT_RING_SLOTS = 64       # frame_range(40) + CONTINUOUS_FRAME + 1 = 45 resident frames + 19 appended ahead

_DTYPES = {torch.float32: capi.F32, torch.float16: capi.F16, torch.bfloat16: capi.BF16}


def sample_frames(frame_idx: int, take_range: int, num_refs: int) -> List[int]:
    """src/model/predict.py:74-89, computed by the C library (host arithmetic, no GPU needed).
    Raises ValueError where the reference's np.linspace does (num_refs < 3 and frame_idx > num_refs)."""
    buf = (C.c_int32 * capi.MAX_REFS)()
    n = capi.lib().vosprop_sample_frames(frame_idx, take_range, num_refs, buf)
    if n == capi.ERR_INVALID:
        raise ValueError(capi.lib().vosprop_last_error().decode())
    capi.check(n)
    return list(buf[:n])


def plan_refs(frame_idx: int, take_range: int, num_refs: int, sigma_dense: float, sigma_sparse: float,
              probability_propagation: bool):
    """(ref_frames, ref_sigmas) for one step: sample_frames + the sigma rule of predict.py:59-66."""
    st = capi.Step()
    rc = capi.lib().vosprop_plan_step(frame_idx, take_range, num_refs, sigma_dense, sigma_sparse,
                                      int(probability_propagation), C.byref(st))
    if rc == capi.ERR_INVALID:
        raise ValueError(capi.lib().vosprop_last_error().decode())
    capi.check(rc)
    return list(st.ref_frames[:st.n_refs]), list(st.ref_sigma[:st.n_refs])


def precision_for(dtype: torch.dtype) -> int:
    """Storage mode that keeps embeddings of `dtype` exact: 16-bit floats go through the tensor cores in one
    pass (their products are exact in the fp32 accumulator); fp32 needs the bf16 hi+lo split."""
    return {torch.float16: capi.PREC_F16, torch.bfloat16: capi.PREC_BF16}.get(dtype, capi.PREC_SPLIT3)


IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)     # src/utils/datasets.py:128-131


def normalize_frames(rgb: torch.Tensor, dtype: torch.dtype = torch.float16, mean=IMAGENET_MEAN, std=IMAGENET_STD) -> torch.Tensor:
    """Decoded frames (n,H,W,3) uint8 on the GPU -> (n,3,H,W) normalised tensor in channels-last layout (the element
    order does not change), fp32 (bit-identical to torchvision's ToTensor + Normalize) or fp16 (that value rounded)."""
    if rgb.dtype != torch.uint8 or not rgb.is_cuda or rgb.dim() != 4 or rgb.shape[3] != 3:
        raise TypeError(f'expected a CUDA uint8 (n,H,W,3) tensor, got {rgb.dtype} {tuple(rgb.shape)} on {rgb.device}')
    if dtype not in (torch.float32, torch.float16):
        raise TypeError('normalize_frames writes fp32 or fp16')
    rgb = rgb.contiguous()
    n, H, W, _ = rgb.shape
    out = torch.empty((n, 3, H, W), dtype=dtype, device=rgb.device, memory_format=torch.channels_last)
    m3, s3 = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    with torch.cuda.device(rgb.device):
        capi.check(capi.lib().vosprop_normalize_u8(C.c_void_p(rgb.data_ptr()), n * H * W, m3, s3, C.c_void_p(out.data_ptr()),
                                                    _DTYPES[dtype], C.c_void_p(torch.cuda.current_stream(rgb.device).cuda_stream)))
    rgb.record_stream(torch.cuda.current_stream(rgb.device))
    return out


def required_ring_slots(frame_range: int, ref_num: int) -> int:
    """Slots needed so that every frame sample_frames can pick is still resident, plus the target."""
    return max(frame_range + CONTINUOUS_FRAME, ref_num) + 1


def _on_device(fn):
    """Run a method with the engine's device current: the library launches on the current CUDA device, and an engine built
    for cuda:1 may be driven from a process whose current device is cuda:0."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *a, **k):
        if torch.cuda.current_device() == (self.device.index or 0):
            return fn(self, *a, **k)
        with torch.cuda.device(self.device):
            return fn(self, *a, **k)
    return wrapper


class PropagationEngine:
    def __init__(self, max_pixels: int, ring_slots: int = DEFAULT_RING_SLOTS, max_fullres_pixels: int = 0,
                 device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise RuntimeError('PropagationEngine needs a CUDA device (sm_100a); there is no CPU fallback')
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self._lib = capi.lib()
        self._h = C.c_void_p()
        cfg = capi.Config(self.device.index or 0, int(max_pixels), int(ring_slots), int(max_fullres_pixels))
        with torch.cuda.device(self.device):
            capi.check(self._lib.vosprop_create(C.byref(cfg), C.byref(self._h)))
        self.ring_slots = ring_slots
        self.max_pixels = max_pixels
        self.geom = None

    def close(self):
        if getattr(self, '_h', None) and self._h.value:
            self._lib.vosprop_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self) -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def num_sms(self) -> int:
        return self._lib.vosprop_num_sms(self._h)

    @property
    def launch_count(self) -> int:
        return self._lib.vosprop_launch_count(self._h)

    @_on_device
    def block_skip(self, mode='auto'):
        """Skipping of blocks whose soft-max weight is below fp32 underflow (vos_prop.h: vosprop_block_skip):
        False / 'off', True / 'on', or 'auto' (the engine's default: probes, then follows what the launches report)."""
        code = {False: 0, True: 1, 'off': 0, 'on': 1, 'auto': 2, 0: 0, 1: 1, 2: 2}[mode]
        capi.check(self._lib.vosprop_block_skip(self._h, code))

    @property
    def block_skip_active(self) -> bool:
        """Would the next launch of the fused index kernel skip dead blocks (auto mode: what the reports so far say)?"""
        return bool(self._lib.vosprop_block_skip_state(self._h))

    @_on_device
    def enable_timing(self, capacity: int, classes=('append', 'affinity', 'merge')):
        """Bracket the kernel launches of the given classes with CUDA events (bench.py roofline); 0 disables."""
        mask = sum(1 << ('append', 'affinity', 'merge').index(c) for c in classes)
        capi.check(self._lib.vosprop_timing_select(self._h, mask))
        capi.check(self._lib.vosprop_timing_enable(self._h, int(capacity)))

    @_on_device
    def read_timing(self):
        """{'append'|'affinity'|'merge': (total_ms, launches)} since the last read (synchronises)."""
        tot, cnt = (C.c_double * 3)(), (C.c_int64 * 3)()
        capi.check(self._lib.vosprop_timing_read(self._h, tot, cnt))
        return {k: (tot[i], cnt[i]) for i, k in enumerate(('append', 'affinity', 'merge'))}

    # ------------------------------------------------------------------ per-video state
    @_on_device
    def reset(self, H_d: int, W_d: int, H: int, W: int, d: int, precision: int = capi.PREC_SPLIT3):
        """New video.  `precision`: PREC_SPLIT3 (any embedding dtype, bf16 hi+lo, 3 tensor-core passes) or
        PREC_F16 / PREC_BF16 (embeddings already 16-bit: one exact pass); see precision_for()."""
        capi.check(self._lib.vosprop_reset(self._h, H_d, W_d, H, W, d, int(precision), self._stream()))
        self.geom = (H_d, W_d, H, W, d)
        self.precision = int(precision)

    @_on_device
    def append(self, frame_idx: int, features: torch.Tensor):
        """features: (K,H_d,W_d) or (1,K,H_d,W_d), fp32/fp16/bf16, standard or channels_last."""
        if features.dim() == 4:
            if features.shape[0] != 1:
                raise ValueError('append() takes one frame')
            features = features[0]
        H_d, W_d = self.geom[0], self.geom[1]
        if tuple(features.shape) != (capi.FEAT_DIM, H_d, W_d):
            raise ValueError(f'expected ({capi.FEAT_DIM},{H_d},{W_d}) features, got {tuple(features.shape)}')
        if features.dtype not in _DTYPES or not features.is_cuda:
            raise TypeError(f'features must be a CUDA fp32/fp16/bf16 tensor, got {features.dtype} on {features.device}')
        if features.is_contiguous():
            layout = capi.NCHW
        elif features.permute(1, 2, 0).is_contiguous():
            layout = capi.NHWC
        else:
            features, layout = features.contiguous(), capi.NCHW
        capi.check(self._lib.vosprop_append_features(self._h, frame_idx, C.c_void_p(features.data_ptr()),
                                                     _DTYPES[features.dtype], layout, self._stream()))
        features.record_stream(torch.cuda.current_stream(self.device))

    def lookahead(self, frame_range: int, ref_num: int) -> int:
        """How many frames may sit in the ring ahead of the frame being propagated (the target included) without
        overwriting a frame sample_frames can still pick: ring_slots - required_ring_slots + 1 (>= 1)."""
        return max(1, self.ring_slots - required_ring_slots(frame_range, ref_num) + 1)

    @_on_device
    def append_frames(self, first_frame_idx: int, features: torch.Tensor, class_idx: Optional[torch.Tensor] = None):
        """features: (n,K,H_d,W_d) fp32/fp16/bf16, standard or channels_last -> frames first_frame_idx .. +n-1 in ONE
        launch; class_idx: optional (n,H_d,W_d) / (n,P) uint8 index labels for all of them (one call installs a labelled clip)."""
        H_d, W_d = self.geom[0], self.geom[1]
        n = features.shape[0]
        if tuple(features.shape[1:]) != (capi.FEAT_DIM, H_d, W_d):
            raise ValueError(f'expected (n,{capi.FEAT_DIM},{H_d},{W_d}) features, got {tuple(features.shape)}')
        if features.dtype not in _DTYPES or not features.is_cuda:
            raise TypeError(f'features must be a CUDA fp32/fp16/bf16 tensor, got {features.dtype} on {features.device}')
        if features.is_contiguous():
            layout = capi.NCHW
        elif features.permute(0, 2, 3, 1).is_contiguous():
            layout = capi.NHWC
        else:
            features, layout = features.contiguous(), capi.NCHW
        cls_ptr = None
        if class_idx is not None:
            class_idx = class_idx.reshape(n, -1).to(device=self.device, dtype=torch.uint8).contiguous()
            if class_idx.shape[1] != H_d * W_d:
                raise ValueError(f'expected {H_d * W_d} labels per frame, got {class_idx.shape[1]}')
            cls_ptr = C.c_void_p(class_idx.data_ptr())
        capi.check(self._lib.vosprop_append_frames(self._h, first_frame_idx, n, C.c_void_p(features.data_ptr()),
                                                   _DTYPES[features.dtype], layout, cls_ptr, self._stream()))
        features.record_stream(torch.cuda.current_stream(self.device))
        if class_idx is not None:
            class_idx.record_stream(torch.cuda.current_stream(self.device))

    @_on_device
    def set_labels_index(self, frame_idx: int, class_idx: torch.Tensor):
        P = self.geom[0] * self.geom[1]
        class_idx = class_idx.reshape(-1).to(device=self.device, dtype=torch.uint8).contiguous()
        if class_idx.numel() != P:
            raise ValueError(f'expected {P} labels, got {class_idx.numel()}')
        capi.check(self._lib.vosprop_set_labels_index(self._h, frame_idx, C.c_void_p(class_idx.data_ptr()), self._stream()))
        class_idx.record_stream(torch.cuda.current_stream(self.device))

    @_on_device
    def set_labels_dense(self, frame_idx: int, labels: torch.Tensor):
        P, d = self.geom[0] * self.geom[1], self.geom[4]
        labels = labels.reshape(labels.shape[0], -1).to(device=self.device, dtype=torch.float32).contiguous()
        if tuple(labels.shape) != (d, P):
            raise ValueError(f'expected ({d},{P}) labels, got {tuple(labels.shape)}')
        capi.check(self._lib.vosprop_set_labels_dense(self._h, frame_idx, C.c_void_p(labels.data_ptr()), self._stream()))
        labels.record_stream(torch.cuda.current_stream(self.device))

    # ------------------------------------------------------------------ the hot path
    @_on_device
    def propagate(self, frame_idx: int, ref_frames: Sequence[int], ref_sigmas: Sequence[float],
                  temperature: float = 1.0, probability_propagation: bool = False, write_labels: bool = True,
                  kernel: int = capi.KERNEL_TC, want_prediction: bool = True, want_lowres: bool = True,
                  want_fullres: bool = True, out_fullres: Optional[torch.Tensor] = None,
                  out_prediction: Optional[torch.Tensor] = None, topk: int = 0,
                  want_topk_idx: bool = False, wait_event: Optional[torch.cuda.Event] = None,
                  record_event: Optional[torch.cuda.Event] = None) -> Dict[str, torch.Tensor]:
        """One propagation step.  topk = 0 is the reference (softmax over every reference pixel); topk = k in
        1..MAX_TOPK is the top-k extension (softmax over the k largest logits per target pixel), optionally
        returning the (P, k) reference indices r*P + pixel, best first."""
        H_d, W_d, H, W, d = self.geom
        P = H_d * W_d
        st = capi.Step()
        st.frame_idx = frame_idx
        st.n_refs = len(ref_frames)
        if len(ref_frames) > capi.MAX_REFS or len(ref_frames) != len(ref_sigmas):
            raise ValueError('bad reference list')
        for i, (f, s) in enumerate(zip(ref_frames, ref_sigmas)):
            st.ref_frames[i] = int(f)
            st.ref_sigma[i] = float(s) if s is not None else 0.0
        st.temperature = float(temperature)
        st.probability_propagation = int(probability_propagation)
        st.write_labels = int(write_labels)
        st.topk = int(topk)
        st.kernel = int(kernel)
        out: Dict[str, torch.Tensor] = {}
        if want_prediction or out_prediction is not None:
            pred = out_prediction if out_prediction is not None else torch.empty((d, P), dtype=torch.float32, device=self.device)
            assert pred.is_contiguous() and pred.dtype == torch.float32 and pred.numel() == d * P
            st.out_prediction = pred.data_ptr()
            out['prediction'] = pred
        if want_lowres:
            low = torch.empty((P,), dtype=torch.uint8, device=self.device)
            st.out_mask_lowres = low.data_ptr()
            out['mask_lowres'] = low
        if want_fullres or out_fullres is not None:
            full = out_fullres if out_fullres is not None else torch.empty((H, W), dtype=torch.uint8, device=self.device)
            assert full.is_contiguous() and full.dtype == torch.uint8 and full.numel() == H * W
            st.out_mask_fullres = full.data_ptr()
            out['mask'] = full
        if want_topk_idx:
            if topk <= 0:
                raise ValueError('want_topk_idx needs topk > 0')
            tk = torch.empty((P, topk), dtype=torch.int32, device=self.device)
            st.out_topk_idx = tk.data_ptr()
            out['topk_idx'] = tk
        if wait_event is not None:      # chain the affinity kernels of several sequences in flight (vos_prop.h)
            st.wait_event = wait_event.cuda_event
        if record_event is not None:
            st.record_event = record_event.cuda_event
        capi.check(self._lib.vosprop_propagate(self._h, C.byref(st), self._stream()))
        return out

    def step(self, frame_idx: int, frame_range: int, ref_num: int, sigma_1: float, sigma_2: float,
             temperature: float, probability_propagation: bool, **kw) -> Dict[str, torch.Tensor]:
        """One call of the reference's predict() + write-back for target `frame_idx`."""
        refs, sig = plan_refs(frame_idx, frame_range, ref_num, sigma_1, sigma_2, probability_propagation)
        return self.propagate(frame_idx, refs, sig, temperature, probability_propagation, **kw)
