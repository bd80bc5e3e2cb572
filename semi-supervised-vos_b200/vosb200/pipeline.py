"""ClipSegmenter: the library-level public call -- frames + first-frame annotation in, masks out.

End-to-end path of one clip on one GPU:
  pinned host frames (decoded uint8 RGB, or already normalised fp32) --H2D (copy stream, batch ahead)--> normalisation on the
  GPU for uint8 input (vosprop_normalize_u8: 4x less PCIe traffic than fp32 frames) --> VOSNet on cuDNN (fp16 autocast, as the
  reference does on CUDA: inference_utils.py:35,52; channels_last) --> per frame: ring append +
  fused propagation (libvosprop) --> uint8 masks accumulate on the device --> one D2H per clip (its own stream: the copy
  runs behind the next clip's compute).
Feature extraction does not depend on propagation state, so frames go through
the backbone in batches while propagation stays strictly sequential.  No host sync inside a clip.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _capi as capi
from .engine import PropagationEngine, normalize_frames, required_ring_slots
from .sequence import start_sequence


class ClipSegmenter:
    def __init__(self, model: torch.nn.Module, device=None, sigma_1: float = 8.0, sigma_2: float = 21.0,
                 frame_range: int = 40, ref_num: int = 9, temperature: float = 1.0,
                 probability_propagation: bool = False, backbone_batch: int = 35, amp: bool = True,
                 channels_last: bool = True, kernel: int = capi.KERNEL_TC, fuse_backbone: bool = True):
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.model = model.to(self.device).eval()
        if channels_last:
            self.model = self.model.to(memory_format=torch.channels_last)
        self.channels_last = channels_last
        self.amp = amp
        # cuDNN-fused inference form (BN folded, bias/ReLU/residual in the conv epilogue) when available
        self.fused = None
        if fuse_backbone and amp and getattr(model, 'model', None) in ('resnet18', 'resnet50', 'resnet101'):
            from .fused_backbone import FusedVOSNet
            self.fused = FusedVOSNet(self.model)
        self.params = dict(sigma_1=sigma_1, sigma_2=sigma_2, frame_range=frame_range, ref_num=ref_num,
                           temperature=temperature, probability_propagation=probability_propagation)
        self.backbone_batch = backbone_batch
        self.kernel = kernel
        self.engine: Optional[PropagationEngine] = None
        self.copy_stream = torch.cuda.Stream(self.device)
        self.d2h_stream = torch.cuda.Stream(self.device)      # masks -> host; apart from the H2D stream: the two directions overlap
        self._d2h_done: Optional[torch.cuda.Event] = None

    def _ensure_engine(self, n_pixels: int):
        slots = required_ring_slots(self.params['frame_range'], self.params['ref_num']) + 19     # 19 frames appended ahead
        if self.engine is None or self.engine.max_pixels < n_pixels or self.engine.ring_slots < slots:
            if self.engine is not None:
                self.engine.close()
            self.engine = PropagationEngine(max_pixels=n_pixels, ring_slots=slots, device=self.device)

    @torch.no_grad()
    def embed(self, frames: torch.Tensor) -> torch.Tensor:
        if self.fused is not None:
            return self.fused(frames)
        x = frames.contiguous(memory_format=torch.channels_last) if self.channels_last else frames
        with torch.autocast('cuda', dtype=torch.float16, enabled=self.amp):
            return self.model(x)

    @torch.no_grad()
    def segment(self, frames: torch.Tensor, first_label: torch.Tensor, out: Optional[torch.Tensor] = None,
                sync: bool = True) -> torch.Tensor:
        """frames: (T,H,W,3) uint8 decoded RGB (normalised on the GPU like datasets.py:128-131) or (T,3,H,W) fp32 already
        normalised; pinned host or device.  first_label (H,W) integer class map.
        Returns masks for frames 1..T-1 as (T-1,H,W) uint8 in pinned host memory.  sync = False: the copy to the host may
        still be in flight on return (wait() or a device synchronise completes it)."""
        raw = frames.dtype == torch.uint8
        if raw:
            T, H, W, _ = frames.shape
        else:
            T, _, H, W = frames.shape
        main = torch.cuda.current_stream(self.device)
        B = self.backbone_batch
        on_host = not frames.is_cuda
        staged = {}

        def stage(b0):
            if not on_host:
                return frames[b0:b0 + B]
            with torch.cuda.stream(self.copy_stream):
                dev = frames[b0:b0 + B].to(self.device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
            staged[b0] = ev
            return dev

        masks_dev = torch.empty((T - 1, H, W), dtype=torch.uint8, device=self.device)
        nxt = stage(0)
        # class count of the clip (predict.py:113) from the host copy of the annotation when there is one: .item() on a device
        # tensor would stall the host on everything queued so far, once per clip
        d = int(first_label.max()) + 1 if not first_label.is_cuda else None
        p = self.params
        for b0 in range(0, T, B):
            cur = nxt
            if b0 + B < T:
                nxt = stage(b0 + B)          # copy of the next batch overlaps this batch's compute
            if on_host:
                main.wait_event(staged.pop(b0))
                cur.record_stream(main)
            if raw:
                cur = normalize_frames(cur, torch.float16 if self.amp else torch.float32)
            feats = self.embed(cur)
            appended = 0                     # frames of this batch already in the ring
            for i in range(feats.shape[0]):
                t = b0 + i
                if t == 0:
                    self._ensure_engine(feats.shape[2] * feats.shape[3])
                    start_sequence(self.engine, feats[0], first_label.to(self.device, non_blocking=True), d=d)
                    appended = 1
                    continue
                if i >= appended:            # one launch appends the next frames of the batch (engine.lookahead)
                    n = min(self.engine.lookahead(p['frame_range'], p['ref_num']), feats.shape[0] - appended)
                    self.engine.append_frames(t, feats[appended:appended + n])
                    appended += n
                self.engine.step(t, p['frame_range'], p['ref_num'], p['sigma_1'], p['sigma_2'], p['temperature'],
                                 p['probability_propagation'], kernel=self.kernel, want_prediction=False,
                                 want_lowres=False, want_fullres=False, out_fullres=masks_dev[t - 1])
        if out is None:
            out = torch.empty((T - 1, H, W), dtype=torch.uint8, pin_memory=True)
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(self.d2h_stream):
            self.d2h_stream.wait_event(ready)
            out.copy_(masks_dev, non_blocking=True)
            masks_dev.record_stream(self.d2h_stream)
            self._d2h_done = torch.cuda.Event()
            self._d2h_done.record(self.d2h_stream)
        if sync:
            self.wait()
        return out

    def wait(self):
        """Blocks until the masks of the last segment(..., sync=False) call are in host memory."""
        if self._d2h_done is not None:
            self._d2h_done.synchronize()
            self._d2h_done = None


class ClipSegmenterPool:
    """Several clips in flight on one GPU: `lanes` ClipSegmenters (each its own engine = reference memory, and its own
    stream) take the clips of a list round-robin.

    Frames of one clip are strictly serial, and both its stages fill the GPU unevenly: the fused launch of a frame ends when
    its slowest CTA does (the others idle for up to 20 us), the merge kernel behind it occupies the SMs
    for 10 us with a few hundred threads each, and the backbone's layers have tails of their own.  A second clip's kernels run
    in those holes.  Measured at 480p (tools/e2e_lanes_probe.py): 2 lanes +2.5 % frames/s end to end, 3 lanes +3.2 %; the masks are
    bit-identical to one clip at a time (every clip still sees exactly its own state)."""

    def __init__(self, model: torch.nn.Module, lanes: int = 2, device=None, **segmenter_args):
        if lanes < 1:
            raise ValueError('lanes >= 1')
        self.segmenters = [ClipSegmenter(model, device=device, **segmenter_args) for _ in range(lanes)]
        self.device = self.segmenters[0].device
        self.streams = [torch.cuda.Stream(self.device) for _ in range(lanes)]

    @torch.no_grad()
    def segment_many(self, clips, outs=None, sync: bool = True):
        """clips: [(frames, first_label), ...] as for ClipSegmenter.segment; outs: optional pinned (T-1,H,W) uint8 tensors.
        Returns the list of mask tensors (pinned host memory).  sync = False: the copies may still be in flight (wait())."""
        main = torch.cuda.current_stream(self.device)
        for s in self.streams:
            s.wait_stream(main)
        results = []
        n = len(self.segmenters)
        for i, (frames, first_label) in enumerate(clips):
            with torch.cuda.stream(self.streams[i % n]):
                results.append(self.segmenters[i % n].segment(frames, first_label, out=None if outs is None else outs[i], sync=False))
        for s in self.streams:
            main.wait_stream(s)
        if sync:
            self.wait()
        return results

    def wait(self):
        for seg in self.segmenters:
            seg.wait()

    def close(self):
        for seg in self.segmenters:
            if seg.engine is not None:
                seg.engine.close()
                seg.engine = None
