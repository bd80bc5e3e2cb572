"""Small helpers with the reference's names (src/utils/utils.py): one-hot, checkpoint loading,
palette-PNG output."""
import os
from pathlib import Path

import numpy as np
import torch
from loguru import logger
from PIL import Image

from src.config import Config


def index_to_onehot(idx, d):
    """(n,) class indices -> (d, n) fp32 one-hot on Config.DEVICE (utils.py:59-68)."""
    idx = idx.reshape(1, -1).long()
    return torch.zeros(d, idx.shape[1], device=Config.DEVICE).scatter_(0, idx.to(Config.DEVICE), 1)


def color_to_class(img, centroids):
    """RGB annotation (B,3,H,W) float -> (B,H,W) long index of the nearest centroid (utils.py:45-56; Euclidean, the
    first minimum wins).  The squared distances are small integers, exact in fp32, so skipping the reference's sqrt
    cannot change an arg-min."""
    B, C, H, W = img.shape
    px = img.permute(0, 2, 3, 1).reshape(-1, 1, C)
    d2 = ((px - centroids.to(px.device).reshape(1, -1, C)) ** 2).sum(2)
    return d2.argmin(1).reshape(B, H, W)


def annotation_centroids():
    """The 22 annotation colours the reference ships as ./annotation_centroids.npy (loaded from the working directory,
    validation.py:75): the first 22 DAVIS palette entries with 192 at 191 (k-means output).  Used when that file is not
    in the working directory."""
    table = np.zeros((22, 3), dtype=np.int32)
    for k in range(22):
        for bit in range(8):
            for ch in range(3):
                table[k, ch] |= ((k >> (3 * bit + ch)) & 1) << (7 - bit)
    table[table == 192] = 191
    return table


def _read_checkpoint(model, checkpoint):
    if checkpoint is None:
        return model
    if not os.path.isfile(checkpoint):
        logger.info("=> no checkpoint found at '{}'".format(checkpoint))
        exit(-1)   # reference behaviour (utils.py:83-85)
    logger.info("=> loading checkpoint '{}'".format(checkpoint))
    blob = torch.load(checkpoint, map_location=Config.DEVICE)
    state = blob['state_dict'] if isinstance(blob, dict) and 'state_dict' in blob else blob
    model.load_state_dict(state)
    logger.info("=> loaded checkpoint '{}'".format(checkpoint))
    return model


def load_model(model, checkpoint):
    """Accepts {'state_dict': ...} or a bare state-dict; retries through DataParallel for
    `module.`-prefixed keys (utils.py:71-94)."""
    try:
        return _read_checkpoint(model, checkpoint)
    except Exception:  # noqa: BLE001 - the reference retries on any failure
        wrapped = torch.nn.DataParallel(model)
        return _read_checkpoint(wrapped, checkpoint).module


def save_prediction(prediction, palette, save_path, save_name, video_name):
    """(H,W) class indices -> `<save_path>/<video_name>/<save_name>.png`, mode P with the
    annotation's palette (utils.py:34-42)."""
    img = Image.fromarray(np.asarray(prediction).astype(np.uint8), mode='L')
    img.putpalette(palette)
    img = img.convert('P')
    video_path = Path(save_path) / video_name
    video_path.mkdir(parents=True, exist_ok=True)
    img.save((video_path / (save_name + '.png')).absolute())


def save_predictions(predictions, palette, save, video_name):
    """frames 1..T-1 -> 00001.png ... (utils.py:97-100)"""
    for idx, prediction in enumerate(predictions, start=1):
        save_prediction(prediction, palette, save, str(idx).zfill(5), video_name)
