"""Per-sequence propagation loops (reference: src/utils/inference_utils.py).

`inference_single` keeps the reference's signature and semantics -- iterate frames, switch
sequence when the video name changes, frame 0 installs the ground truth, every later frame is
predicted from the reference memory, results are saved as palette PNGs -- but the state lives in
a PropagationEngine (bounded ring on the GPU) instead of ever-growing torch.cat histories, and
nothing synchronises per frame: masks accumulate in a device buffer and cross to the host once
per video.
"""
from __future__ import annotations

import numpy as np
import torch
from tqdm import tqdm

from src.config import Config
from src.model.predict import prepare_first_frame
from src.utils.utils import save_predictions
from vosb200 import PropagationEngine
from vosb200.engine import precision_for, required_ring_slots

REDUCTIONS = {'maximum': lambda x, y: torch.maximum(x, y),
              'minimum': lambda x, y: torch.minimum(x, y),
              'mean': lambda x, y: (x + y) / 2.0}

_ENGINES = {}


def _engine_for(n_pixels: int, slots: int) -> PropagationEngine:
    dev = torch.device(Config.DEVICE)
    if dev.type != 'cuda':
        raise RuntimeError("this build runs the propagation on a B200 only: use --device cuda "
                           "(there is no CPU path; the CPU reference lives in oracle/ for tests)")
    key = (dev.index or 0,)
    eng = _ENGINES.get(key)
    if eng is None or eng.max_pixels < n_pixels or eng.ring_slots < slots:
        if eng is not None:
            eng.close()
        eng = _ENGINES[key] = PropagationEngine(max_pixels=n_pixels, ring_slots=max(slots, 48), device=dev)
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

    def flush(self):
        if not self.chunks:
            return
        parts = [c for c in self.chunks[:-1]] + [self.chunks[-1][:self.fill]]
        masks = torch.cat(parts, 0).cpu().numpy()   # the only sync of the video
        save_predictions(masks, self.palette, self.save, self.video)


def inference_single(model, inference_loader, total_len, annotation_dir, last_video, save, sigma_1, sigma_2,
                     frame_range, ref_num, temperature, probability_propagation, disable):
    """Reference: src/utils/inference_utils.py:23-87."""
    frame_idx = 0
    sink = None
    engine = None
    slots = required_ring_slots(frame_range, ref_num)
    for input, (current_video,) in tqdm(inference_loader, total=total_len, disable=disable):
        if current_video != last_video:
            if sink is not None:
                sink.flush()
                sink = None
            frame_idx = 0
        input = input.to(Config.DEVICE, non_blocking=True)
        with torch.autocast('cuda', dtype=torch.float16):
            features = model(input)
        if frame_idx == 0:
            first_annotation = annotation_dir / current_video / '00000.png'
            label_1hot, d, palette, _, _ = prepare_first_frame(
                current_video, save, first_annotation, sigma_1, sigma_2, inference_strategy='single',
                probability_propagation=probability_propagation)
            (_, _, H, W) = input.shape
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


def _not_built(name):
    def strategy(*args, **kwargs):
        raise NotImplementedError(
            f"inference strategy '{name}' (test-time augmentation, reference src/utils/inference_utils.py) is "
            f"outside this round's hot-path scope (SURVEY.md section 8f, row N2); use --inference-strategy single")
    strategy.__name__ = name
    return strategy


inference_hor_flip = _not_built('inference_hor_flip')
inference_ver_flip = _not_built('inference_ver_flip')
inference_2_scale = _not_built('inference_2_scale')
inference_multimodel = _not_built('inference_multimodel')
inference_3_scale = _not_built('inference_3_scale')
