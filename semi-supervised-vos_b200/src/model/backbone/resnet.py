"""Stride-8 ResNet trunk for VOSNet (stays on cuDNN -- BASELINE.json north_star).

Architecture facts taken from the reference (src/model/backbone/resnet.py:99-156): torchvision-style
stem (7x7/2 conv, BN, ReLU, 3x3/2 max-pool), layer1 (stride 1), layer2 (stride 2), layer3 and layer4
at stride 1, and layer4 built with planes=256 (so a Bottleneck trunk ends with 1024 channels,
a BasicBlock trunk with 256... x expansion).  Module / parameter names are kept (conv1, bn1,
layer1..4, <block>.conv{1,2,3}, bn{1,2,3}, downsample.{0,1}) so the reference's checkpoints load
(state-dict keys `backbone.{0,1,4,5,6,7}.*` once wrapped by VOSNet).  The classifier head
(avgpool/fc) is kept only so that `list(children())[0:8]` slices the same eight trunk modules.
"""
import math

import torch.nn as nn
import torch.utils.model_zoo as model_zoo

IMAGENET_URLS = {
    'resnet18': 'https://download.pytorch.org/models/resnet18-5c106cde.pth',
    'resnet50': 'https://download.pytorch.org/models/resnet50-19c8e357.pth',
    'resnet101': 'https://download.pytorch.org/models/resnet101-5d3b4d8f.pth',
}


def _conv_bn(cin, cout, k, stride, norm):
    return [nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=k // 2, bias=False), norm(cout)]


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, BatchNorm=nn.BatchNorm2d):
        super().__init__()
        self.conv1, self.bn1 = _conv_bn(inplanes, planes, 3, stride, BatchNorm)
        self.relu = nn.ReLU(inplace=True)
        self.conv2, self.bn2 = _conv_bn(planes, planes, 3, 1, BatchNorm)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        skip = x if self.downsample is None else self.downsample(x)
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return self.relu(y + skip)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, BatchNorm=nn.BatchNorm2d):
        super().__init__()
        self.conv1, self.bn1 = _conv_bn(inplanes, planes, 1, 1, BatchNorm)
        self.conv2, self.bn2 = _conv_bn(planes, planes, 3, stride, BatchNorm)   # stride sits on the 3x3
        self.conv3, self.bn3 = _conv_bn(planes, planes * 4, 1, 1, BatchNorm)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        skip = x if self.downsample is None else self.downsample(x)
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        return self.relu(y + skip)


class ResNet(nn.Module):
    # (planes, stride) per stage: the last two stages keep stride 1 and layer4 uses 256 planes
    STAGES = ((64, 1), (128, 2), (256, 1), (256, 1))

    def __init__(self, block, layers, BatchNorm=nn.BatchNorm2d, num_classes=1000):
        super().__init__()
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = BatchNorm(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        for i, ((planes, stride), n) in enumerate(zip(self.STAGES, layers), start=1):
            setattr(self, f'layer{i}', self._stage(block, planes, n, stride, BatchNorm))
        self.avgpool = nn.AvgPool2d(7, stride=1)
        self.fc = nn.Linear(512 * block.expansion, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):   # He init on fan-out, as the reference (resnet.py:116-119)
                fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                nn.init.normal_(m.weight, 0.0, math.sqrt(2.0 / fan_out))
            elif isinstance(m, BatchNorm):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    def _stage(self, block, planes, n_blocks, stride, norm):
        out_ch = planes * block.expansion
        down = None
        if stride != 1 or self.inplanes != out_ch:
            down = nn.Sequential(nn.Conv2d(self.inplanes, out_ch, kernel_size=1, stride=stride, bias=False), norm(out_ch))
        blocks = [block(self.inplanes, planes, stride, down, BatchNorm=norm)]
        self.inplanes = out_ch
        blocks += [block(out_ch, planes, BatchNorm=norm) for _ in range(n_blocks - 1)]
        return nn.Sequential(*blocks)

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        x = self.avgpool(x)
        return self.fc(x.view(x.size(0), -1))


def _build(name, block, layers, pretrained, BatchNorm, **kwargs):
    model = ResNet(block, layers, BatchNorm=BatchNorm, **kwargs)
    if pretrained:
        # ImageNet initialisation of conv1..layer3 (layer4/fc differ in shape), as the reference
        # does (resnet.py:192-199).  Offline this cannot be fetched; the --resume checkpoint
        # overwrites every weight anyway, so a failed download only downgrades to a warning.
        try:
            weights = model_zoo.load_url(IMAGENET_URLS[name])
            weights = {k: v for k, v in weights.items() if not k.startswith(('layer4', 'fc'))}
            state = model.state_dict()
            state.update(weights)
            model.load_state_dict(state)
        except Exception as exc:  # noqa: BLE001
            from loguru import logger
            logger.warning(f'ImageNet weights for {name} unavailable ({type(exc).__name__}); keeping random init')
    return model


def resnet18(pretrained=False, BatchNorm=nn.BatchNorm2d, **kwargs):
    return _build('resnet18', BasicBlock, [2, 2, 2, 2], pretrained, BatchNorm, **kwargs)


def resnet50(pretrained=False, BatchNorm=nn.BatchNorm2d, **kwargs):
    return _build('resnet50', Bottleneck, [3, 4, 6, 3], pretrained, BatchNorm, **kwargs)


def resnet101(pretrained=False, BatchNorm=nn.BatchNorm2d, **kwargs):
    return _build('resnet101', Bottleneck, [3, 4, 23, 3], pretrained, BatchNorm, **kwargs)
