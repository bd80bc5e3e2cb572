"""Per-sequence propagation loops (reference: src/utils/inference_utils.py).

`inference_single` keeps the reference's signature and semantics -- iterate frames, switch
sequence when the video name changes, frame 0 installs the ground truth, every later frame is
predicted from the reference memory, results are saved as palette PNGs -- but the state lives in
a PropagationEngine (bounded ring on the GPU) instead of ever-growing torch.cat histories, and
nothing synchronises per frame: masks accumulate in a device buffer and cross to the host once
per video.
"""
from __future__ import annotations

import os

import numpy as np
import torch
from tqdm import tqdm

from src.config import Config
from src.model.predict import prepare_first_frame
from src.utils.utils import save_prediction
from vosb200 import PropagationEngine, normalize_frames
from vosb200.engine import precision_for, required_ring_slots
from vosb200.sequence import nearest_index

REDUCTIONS = {'maximum': lambda x, y: torch.maximum(x, y),
              'minimum': lambda x, y: torch.minimum(x, y),
              'mean': lambda x, y: (x + y) / 2.0}

_ENGINES = {}
# VOS_BLOCK_SKIP: 0 never / 1 always / auto (default).  The fused kernel can leave out blocks of the affinity matrix whose soft-max
# weight is below fp32 underflow (vos_prop.h: vosprop_block_skip): 6-17 % faster propagation on embeddings as peaked as trained
# ones, 6-9 % slower where nothing can be skipped; in auto mode the engine probes and follows what its launches report.
_BLOCK_SKIP = {'0': 'off', '1': 'on'}.get(os.environ.get('VOS_BLOCK_SKIP', 'auto'), 'auto')


def _require_cuda() -> torch.device:
    dev = torch.device(Config.DEVICE)
    if dev.type != 'cuda':
        raise RuntimeError("this build runs the propagation on a B200 only: use --device cuda "
                           "(there is no CPU path; the CPU reference lives in oracle/ for tests)")
    return dev


def _engine_for(n_pixels: int, slots: int) -> PropagationEngine:
    dev = _require_cuda()
    key = (dev.index or 0,)
    eng = _ENGINES.get(key)
    if eng is None or eng.max_pixels < n_pixels or eng.ring_slots < slots:
        if eng is not None:
            eng.close()
        eng = _ENGINES[key] = PropagationEngine(max_pixels=n_pixels, ring_slots=max(slots, 48), device=dev)
        eng.block_skip(_BLOCK_SKIP)
    return eng


class _VideoSink:
    """Device-side accumulator of one video's masks; one D2H + PNG encode at the end."""

    def __init__(self, video, palette, save, H, W, device):
        self.video, self.palette, self.save = video, palette, save
        self.H, self.W, self.device = H, W, device
        self.chunks, self.fill = [], 0

    def next_slot(self) -> torch.Tensor:
        if not self.chunks or self.fill == self.chunks[-1].shape[0]:
            self.chunks.append(torch.empty((32, self.H, self.W), dtype=torch.uint8, device=self.device))
            self.fill = 0
        self.fill += 1
        return self.chunks[-1][self.fill - 1]

    def stacked(self) -> torch.Tensor:
        """(frames, H, W) uint8 on the device."""
        return torch.cat([c for c in self.chunks[:-1]] + [self.chunks[-1][:self.fill]], 0)

    def flush(self, dev_masks=None):
        """One D2H copy into pinned memory, then PNG encoding on a writer thread: the next video's propagation does
        not wait for the files of this one (reference: save_predictions on the critical path, inference_utils.py:30)."""
        if not self.chunks:
            return
        if dev_masks is None:
            dev_masks = self.stacked()
        host = torch.empty(dev_masks.shape, dtype=torch.uint8, pin_memory=True)
        host.copy_(dev_masks, non_blocking=True)
        done = torch.cuda.Event()
        done.record()
        _WRITER.submit(done, host, self.palette, self.save, self.video)


class _PngWriter:
    """Background writer of finished videos (bounded: at most two videos wait for their files).  One dispatcher thread
    waits for a video's D2H copy and fans its frames out to a small pool: PIL's PNG encoder releases the GIL, and a
    single encoder (about 2-5 ms per 480p mask) would otherwise cap the whole command below 500 frames/s."""

    def __init__(self):
        self._queue, self._thread, self._pool, self._error = None, None, None, None

    def _run(self):
        while True:
            item = self._queue.get()
            try:
                if item is None:
                    return
                done, host, palette, save, video = item
                done.synchronize()
                masks = host.numpy()
                # frames 1..T-1 -> 00001.png ... exactly as save_predictions (utils.py:97-100), several files at a time
                list(self._pool.map(lambda job: save_prediction(job[1], palette, save, str(job[0]).zfill(5), video),
                                    enumerate(masks, start=1)))
            except BaseException as exc:  # noqa: BLE001 - surfaced by drain()
                self._error = exc
            finally:
                self._queue.task_done()

    def submit(self, *item):
        import os
        import queue
        import threading
        from concurrent.futures import ThreadPoolExecutor
        if self._thread is None or not self._thread.is_alive():
            cpus = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
            self._queue, self._error = queue.Queue(maxsize=2), None
            self._pool = ThreadPoolExecutor(max_workers=max(1, min(6, cpus // 2)), thread_name_prefix='vos-png')
            self._thread = threading.Thread(target=self._run, name='vos-png-writer', daemon=True)
            self._thread.start()
        self._queue.put(item)

    def drain(self):
        """Block until every submitted video is on disk; re-raise a writer failure."""
        if self._queue is not None:
            self._queue.join()
            if self._error is not None:
                err, self._error = self._error, None
                raise err


_WRITER = _PngWriter()


def _to_device(input):
    """Loader item -> network input on Config.DEVICE: the reference's normalised fp32 (n,3,H,W) tensor as is, or decoded
    uint8 (n,H,W,3) frames (InferenceDataset(raw=True)) normalised on the GPU -- bit-identical values, a quarter of the
    bytes over PCIe -- or int16 (n,L) JPEG coefficient items (raw='coef') decoded to those frames on the GPU first."""
    if input.dtype == torch.int16:       # InferenceDataset(raw='coef'): Huffman-decoded JPEG coefficients, the GPU finishes the decode
        from vosb200 import jpeg
        return normalize_frames(jpeg.unpack_items(input, _require_cuda()), torch.float32)
    input = input.to(_require_cuda(), non_blocking=True)
    if input.dtype == torch.uint8:
        return normalize_frames(input, torch.float32)
    return input


BACKBONE_LOOKAHEAD = 8     # frames of one video embedded per backbone call (a batch-1 ResNet is launch-bound on a B200)


def amp_enabled() -> bool:
    """The reference runs the network under fp16 autocast on CUDA (inference_utils.py:35,52) and so does this build.
    VOS_AMP=0 is the parity mode against the reference's CPU (fp32) path: fp32 backbone without TF32, embeddings stored as
    bf16 hi + lo (three tensor-core passes) -- tests/test_gpu_e2e_golden.py."""
    import os
    return os.environ.get('VOS_AMP', '1') != '0'


def _embedded(models, inference_loader, total_len, disable, resize=None, shared_input=False):
    """Frame-by-frame view of the loader with the backbone run on up to BACKBONE_LOOKAHEAD consecutive frames of one
    video at a time.  Feature extraction does not depend on the propagation state (the reference calls model(input) on
    every frame independently, inference_utils.py:35,52), so only the launch count changes.
    `models`: one network -> yields (features (1,K,H_d,W_d), (H, W) of the network input, video name); a tuple of
    networks, one per stream of a test-time-augmentation strategy (loader items carry one input per stream, or a single
    input for all of them with shared_input) -> yields (tuple of features, tuple of sizes, video name).
    `resize`: optional (H, W) -> (H', W'), or one per stream, applied with nearest interpolation before the network
    (3-scale)."""
    single = not isinstance(models, (tuple, list))
    nets = (models,) if single else tuple(models)
    pending = []

    def flush():
        feats, sizes = [], []
        base = None
        for k, net in enumerate(nets):
            if base is None or not shared_input:
                base = _to_device(torch.cat([inputs[0 if shared_input else k] for inputs, _ in pending]))
            x, r = base, (resize[k] if isinstance(resize, (tuple, list)) else resize)
            if r is not None:
                x = torch.nn.functional.interpolate(base, size=r(base.shape[2], base.shape[3]), mode='nearest')
            with torch.autocast('cuda', dtype=torch.float16, enabled=amp_enabled()):
                feats.append(net(x))
            sizes.append((x.shape[2], x.shape[3]))
        for i, (_, video) in enumerate(pending):
            if single:
                yield feats[0][i:i + 1], sizes[0], video
            else:
                yield tuple(f[i:i + 1] for f in feats), tuple(sizes), video
        pending.clear()

    for input, (current_video,) in tqdm(inference_loader, total=total_len, disable=disable):
        inputs = (input,) if (single or shared_input) else tuple(input)
        if pending and (current_video != pending[0][1] or any(a.shape != b.shape for a, b in zip(inputs, pending[0][0]))
                        or len(pending) == BACKBONE_LOOKAHEAD):
            yield from flush()
        pending.append((inputs, current_video))
    if pending:
        yield from flush()


def inference_single(model, inference_loader, total_len, annotation_dir, last_video, save, sigma_1, sigma_2,
                     frame_range, ref_num, temperature, probability_propagation, disable):
    """Reference: src/utils/inference_utils.py:23-87."""
    frame_idx = 0
    sink = None
    engine = None
    slots = required_ring_slots(frame_range, ref_num)
    for features, (H, W), current_video in _embedded(model, inference_loader, total_len, disable):
        if current_video != last_video:
            if sink is not None:
                sink.flush()
                sink = None
            frame_idx = 0
        if frame_idx == 0:
            first_annotation = annotation_dir / current_video / '00000.png'
            label_1hot, d, palette, _, _ = prepare_first_frame(
                current_video, save, first_annotation, sigma_1, sigma_2, inference_strategy='single',
                probability_propagation=probability_propagation)
            (_, _, H_d, W_d) = features.shape
            engine = _engine_for(H_d * W_d, slots)
            engine.reset(H_d, W_d, H, W, int(d), precision_for(features.dtype))   # fp16 under autocast -> one exact pass
            engine.append(0, features)
            engine.set_labels_index(0, label_1hot[:, 0].argmax(0))
            sink = _VideoSink(current_video, palette, save, H, W, features.device)
            frame_idx += 1
            last_video = current_video
            continue
        engine.append(frame_idx, features)
        engine.step(frame_idx, frame_range, ref_num, sigma_1, sigma_2, temperature, probability_propagation,
                    want_prediction=False, want_lowres=False, want_fullres=False, out_fullres=sink.next_slot())
        last_video = current_video
        frame_idx += 1
    if sink is not None:
        sink.flush()
    _WRITER.drain()          # every PNG of the run is on disk when the function returns, as in the reference


# ------------------------------------------------------------------------------------------------
# Test-time-augmentation strategies (reference: src/utils/inference_utils.py:90-595).
# Every strategy is "two propagation memories side by side + a per-frame fusion of their outputs"; the
# memories never exchange labels.  Each memory is one PropagationEngine (own ring); the fusion runs on the
# device on the tensors the reference fuses on the host, with the same torch calls (incl. its quirks:
# `torch.fliplr` for BOTH flip strategies, src/utils/inference_utils.py:173,279; `fliplr` of a (1,d,H,W)
# probability map flips the class axis; `.half()` before the arg-max).
# ------------------------------------------------------------------------------------------------
class _Stream:
    """One propagation memory of a multi-stream strategy."""

    def __init__(self, slots):
        self.slots, self.engine = slots, None

    def start(self, features, label_1hot, H, W, d):
        (_, _, H_d, W_d) = features.shape
        n_pixels = H_d * W_d
        if self.engine is None or self.engine.max_pixels < n_pixels:
            if self.engine is not None:
                self.engine.close()
            self.engine = PropagationEngine(max_pixels=n_pixels, ring_slots=max(self.slots, 48), device=features.device)
            self.engine.block_skip(_BLOCK_SKIP)
        self.geom = (H_d, W_d, H, W, int(d))
        self.engine.reset(H_d, W_d, H, W, int(d), precision_for(features.dtype))
        self.engine.append(0, features)
        self.engine.set_labels_index(0, label_1hot[:, 0].argmax(0))

    def step(self, frame_idx, features, p, probability_propagation, want_mask=False):
        """-> (H,W) uint8 label map, or (unless want_mask) the (1,d,H,W) fp32 up-sampled prediction in probability mode."""
        self.engine.append(frame_idx, features)
        want_mask = want_mask or not probability_propagation
        out = self.engine.step(frame_idx, p['frame_range'], p['ref_num'], p['sigma_1'], p['sigma_2'], p['temperature'],
                               probability_propagation, want_prediction=not want_mask, want_lowres=False,
                               want_fullres=want_mask)
        if want_mask:
            return out['mask']
        H_d, W_d, H, W, d = self.geom
        ys, xs = nearest_index(H, H_d, features.device), nearest_index(W, W_d, features.device)
        return out['prediction'].view(d, H_d, W_d)[:, ys][:, :, xs].unsqueeze(0)     # F.interpolate(mode='nearest')


def _fuse(pred_a, pred_b, probability_propagation, reduction_str):
    """inference_utils.py:178-184 (and the same lines of the other strategies) -> (H,W) uint8 on the device."""
    if probability_propagation:
        fused = REDUCTIONS.get(reduction_str)(pred_a, pred_b).half()
        return torch.argmax(fused, 1)[0].to(torch.uint8)
    return torch.maximum(pred_a, pred_b)


def _inference_two_streams(models, inference_loader, total_len, annotation_dir, last_video, save, sigma_1, sigma_2,
                           frame_range, ref_num, temperature, probability_propagation, reduction_str, disable,
                           strategy, scale=None, transform_b=None):
    p = dict(sigma_1=sigma_1, sigma_2=sigma_2, frame_range=frame_range, ref_num=ref_num, temperature=temperature)
    slots = required_ring_slots(frame_range, ref_num)
    a, b = _Stream(slots), _Stream(slots)
    frame_idx, sink = 0, None
    for (features_a, features_b), ((H, W), _), current_video in _embedded(models, inference_loader, total_len, disable,
                                                                       shared_input=strategy == 'multimodel'):
        if current_video != last_video:
            if sink is not None:
                sink.flush()
                sink = None
            frame_idx = 0
        if frame_idx == 0:
            first_annotation = annotation_dir / current_video / '00000.png'
            prepared = prepare_first_frame(current_video, save, first_annotation, sigma_1, sigma_2,
                                           inference_strategy={'vert-flip': 'ver-flip'}.get(strategy, strategy),
                                           probability_propagation=probability_propagation, scale=scale)
            if strategy in ('hor-flip', 'vert-flip'):
                label_a, label_b, d, palette = prepared[0], prepared[1], prepared[2], prepared[3]
            elif strategy in ('2-scale', 'hor-2-scale'):
                (label_a, label_b), d, palette = prepared[0], prepared[1], prepared[2]
            else:   # multimodel: one label set for both models
                label_a = label_b = prepared[0]
                d, palette = prepared[1], prepared[2]
            a.start(features_a, label_a, H, W, d)
            b.start(features_b, label_b, H, W, d)     # both memories predict at the size of the first input
            sink = _VideoSink(current_video, palette, save, H, W, features_a.device)
            frame_idx += 1
            last_video = current_video
            continue
        pred_a = a.step(frame_idx, features_a, p, probability_propagation)
        pred_b = b.step(frame_idx, features_b, p, probability_propagation)
        if transform_b is not None:
            pred_b = transform_b(pred_b)
        sink.next_slot().copy_(_fuse(pred_a, pred_b, probability_propagation, reduction_str))
        last_video = current_video
        frame_idx += 1
    if sink is not None:
        sink.flush()
    _WRITER.drain()


def inference_hor_flip(model, inference_loader, total_len, annotation_dir, last_video, save, sigma_1, sigma_2,
                       frame_range, ref_num, temperature, probability_propagation, reduction_str, disable):
    """Reference: src/utils/inference_utils.py:90-192."""
    _inference_two_streams((model, model), inference_loader, total_len, annotation_dir, last_video, save, sigma_1,
                           sigma_2, frame_range, ref_num, temperature, probability_propagation, reduction_str, disable,
                           'hor-flip', transform_b=torch.fliplr)


def inference_ver_flip(model, inference_loader, total_len, annotation_dir, last_video, save, sigma_1, sigma_2,
                       frame_range, ref_num, temperature, probability_propagation, reduction_str, disable):
    """Reference: src/utils/inference_utils.py:195-298 -- which un-flips the vertical stream with `torch.fliplr`
    (:279), i.e. horizontally; kept, because the drop-in must write the files the reference writes."""
    _inference_two_streams((model, model), inference_loader, total_len, annotation_dir, last_video, save, sigma_1,
                           sigma_2, frame_range, ref_num, temperature, probability_propagation, reduction_str, disable,
                           'vert-flip', transform_b=torch.fliplr)


def inference_2_scale(model, inference_loader, total_len, annotation_dir, last_video, save, sigma_1, sigma_2,
                      frame_range, ref_num, temperature, probability_propagation, scale, reduction_str, flip_pred,
                      disable):
    """Reference: src/utils/inference_utils.py:302-410 ('2-scale', and 'hor-2-scale' with flip_pred)."""
    hflip = (lambda t: torch.flip(t, dims=(-1,))) if flip_pred else None      # torchvision hflip: last axis
    _inference_two_streams((model, model), inference_loader, total_len, annotation_dir, last_video, save, sigma_1,
                           sigma_2, frame_range, ref_num, temperature, probability_propagation, reduction_str, disable,
                           'hor-2-scale' if flip_pred else '2-scale', scale=scale, transform_b=hflip)


def inference_multimodel(model, additional_model, inference_loader, total_len, annotation_dir, last_video, save,
                         sigma_1, sigma_2, frame_range, ref_num, temperature, probability_propagation, reduction_str,
                         disable):
    """Reference: src/utils/inference_utils.py:411-511 (two networks, one input, one label set)."""
    _inference_two_streams((model, additional_model), inference_loader, total_len, annotation_dir, last_video, save,
                           sigma_1, sigma_2, frame_range, ref_num, temperature, probability_propagation, reduction_str,
                           disable, 'multimodel')


THREE_SCALE_OUT = (480, 910)     # the reference up-samples every scale's prediction to this size, whatever the input


def inference_3_scale(model, inference_loader, total_len, annotation_dir, last_video, save, sigma_1, sigma_2,
                      frame_range, ref_num, temperature, probability_propagation, scale, disable):
    """Reference: src/utils/inference_utils.py:514-595: the frames nearest-resized by 0.9 / 1.0 / `scale`, one plain
    single-memory propagation per scale, every label map written at 480 x 910 (hard-coded there, :574; kept, because the
    drop-in must write the files the reference writes), the three maps of a frame fused by an element-wise maximum of the
    class indices (:594).  The reference runs the loader three times, one pass per scale; the three propagations never
    exchange anything, so here they run side by side on one pass (three memories, one JPEG decode per frame) and the
    fused map goes straight to the video's sink."""
    H_out, W_out = THREE_SCALE_OUT
    p = dict(sigma_1=sigma_1, sigma_2=sigma_2, frame_range=frame_range, ref_num=ref_num, temperature=temperature)
    slots = required_ring_slots(frame_range, ref_num)
    scales = (0.9, 1.0, scale)
    streams = [_Stream(slots) for _ in scales]
    resize = [lambda H, W, s=s: (int(np.ceil(H * s)), int(np.ceil(W * s))) for s in scales]
    frame_idx, sink, current = 0, None, None
    for features, sizes, current_video in _embedded((model,) * len(scales), inference_loader, total_len, disable,
                                                    resize=resize, shared_input=True):
        if current is not None and current_video != current:
            sink.flush()
            frame_idx = 0
        current = current_video
        if frame_idx == 0:
            first_annotation = annotation_dir / current_video / '00000.png'
            for k, s in enumerate(scales):
                label_1hot, d, palette, _, _ = prepare_first_frame(
                    current_video, save, first_annotation, sigma_1, sigma_2, inference_strategy='3-scale',
                    probability_propagation=probability_propagation, scale=s)
                (_, _, H_d, W_d) = features[k].shape
                if label_1hot.shape[-1] != H_d * W_d:
                    raise ValueError(f'scale {s}: the network maps {sizes[k]} to {(H_d, W_d)} but the annotation is sampled '
                                     f'at {label_1hot.shape[-1]} pixels (the reference fails on such sizes too)')
                streams[k].start(features[k], label_1hot, H_out, W_out, d)
            sink = _VideoSink(current_video, palette, save, H_out, W_out, features[0].device)
            frame_idx += 1
            continue
        masks = [st.step(frame_idx, features[k], p, probability_propagation, want_mask=True) for k, st in enumerate(streams)]
        sink.next_slot().copy_(torch.maximum(torch.maximum(masks[0], masks[1]), masks[2]))
        frame_idx += 1
    if sink is not None:
        sink.flush()
    _WRITER.drain()
