"""Crop helpers of the training/validation dataset (reference src/utils/transforms.py:13-47).  The random draws
come from torch's global generator in the reference's order, so a seeded run samples the same crops."""
import numbers

import torch


def get_crop_params(img_size, output_size):
    """(w, h) image size -> (top, left, height, width) of a uniformly drawn crop; the whole image if it already
    has the requested size (no RNG draw in that case)."""
    w, h = img_size
    th, tw = (output_size, output_size) if isinstance(output_size, numbers.Number) else output_size
    if (w, h) == (tw, th):
        return 0, 0, h, w
    top = torch.randint(low=0, high=h - th, size=(1,)).item()
    left = torch.randint(low=0, high=w - tw, size=(1,)).item()
    return top, left, th, tw


def crop(img, i, j, h, w):
    """PIL crop with (top, left, height, width) arguments."""
    return img.crop((j, i, j + w, i + h))
