"""Inference-time form of VOSNet that stays on cuDNN but lets cuDNN fuse the conv epilogues.

The stock module graph (src/model/vos_net.py, same as the reference) runs every BatchNorm, ReLU and
residual add as a separate HBM-bound kernel: at 480p the trunk reaches only ~220 TFLOP/s.  In eval mode
BatchNorm is an affine map, so it folds into the preceding convolution (w' = w * g/sqrt(var+eps),
b' = beta - mean * g/sqrt(var+eps)); cuDNN's fused ops then apply bias + ReLU (+ residual add) in the
convolution epilogue:  aten::cudnn_convolution_relu / aten::cudnn_convolution_add_relu.
fp16, channels_last -- the precision class the reference itself uses on CUDA (autocast,
src/utils/inference_utils.py:35,52).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


def _fold(conv: nn.Conv2d, bn: Optional[nn.BatchNorm2d]) -> Tuple[torch.Tensor, torch.Tensor]:
    w = conv.weight.detach().float()
    b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(w.shape[0], device=w.device)
    if bn is not None:
        scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
        w = w * scale.view(-1, 1, 1, 1)
        b = (b - bn.running_mean.detach().float()) * scale + bn.bias.detach().float()
    return w, b


class _Conv:
    """Folded conv (+bias) with optional fused ReLU / residual-add-ReLU."""

    def __init__(self, conv: nn.Conv2d, bn: Optional[nn.BatchNorm2d], dtype: torch.dtype):
        w, b = _fold(conv, bn)
        self.w = w.to(dtype).contiguous(memory_format=torch.channels_last)
        self.b = b.to(dtype)
        self.stride, self.padding, self.dilation, self.groups = conv.stride, conv.padding, conv.dilation, conv.groups

    def plain(self, x):
        return F.conv2d(x, self.w, self.b, self.stride, self.padding, self.dilation, self.groups)

    def relu(self, x):
        return torch.cudnn_convolution_relu(x, self.w, self.b, self.stride, self.padding, self.dilation, self.groups)

    def add_relu(self, x, z):
        return torch.cudnn_convolution_add_relu(x, self.w, z, 1.0, self.b, self.stride, self.padding, self.dilation,
                                                self.groups)


class FusedVOSNet(nn.Module):
    """Built from an eval-mode VOSNet('resnet18'|'resnet50'|'resnet101'); forward(x) -> (B,256,H/8,W/8) fp16,
    channels_last."""

    def __init__(self, net: nn.Module, dtype: torch.dtype = torch.float16):
        super().__init__()
        if net.training:
            raise ValueError('FusedVOSNet folds BatchNorm statistics: put the model in eval() first')
        if net.model not in ('resnet18', 'resnet50', 'resnet101'):
            raise NotImplementedError(f'no fused form for the {net.model} trunk')
        self.dtype = dtype
        bb = net.backbone
        self.stem = _Conv(bb[0], bb[1], dtype)
        self.pool = bb[3]
        self.blocks: List[tuple] = []
        for stage in (bb[4], bb[5], bb[6], bb[7]):
            for blk in stage:
                convs = [_Conv(blk.conv1, blk.bn1, dtype), _Conv(blk.conv2, blk.bn2, dtype)]
                if hasattr(blk, 'conv3'):
                    convs.append(_Conv(blk.conv3, blk.bn3, dtype))
                down = _Conv(blk.downsample[0], blk.downsample[1], dtype) if blk.downsample is not None else None
                self.blocks.append((convs, down))
        self.head = _Conv(net.adjust_dim, net.bn256, dtype) if net.model != 'resnet18' else None

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = x.to(self.dtype).contiguous(memory_format=torch.channels_last)
        x = self.pool(self.stem.relu(x))
        for convs, down in self.blocks:
            skip = x if down is None else down.plain(x)
            y = x
            for c in convs[:-1]:
                y = c.relu(y)
            x = convs[-1].add_relu(y, skip)
        return x if self.head is None else self.head.plain(x)
