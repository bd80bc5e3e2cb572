"""VOSNet: image -> stride-8, 256-d embedding.  Same constructor / forward / state-dict layout as
the reference's src/model/vos_net.py:9-54; the convolutions stay on cuDNN (north_star).  The
propagation engine accepts the output as is (fp32 or autocast fp16, NCHW or channels_last)."""
import torch
import torch.nn as nn

from src.model.backbone.resnet import resnet18, resnet50, resnet101

_TRUNKS = {'resnet18': resnet18, 'resnet50': resnet50, 'resnet101': resnet101}


class VOSNet(nn.Module):
    def __init__(self, model='resnet50', pretrained=True):
        super().__init__()
        self.model = model
        if model in _TRUNKS:
            trunk = _TRUNKS[model](pretrained=pretrained)
            self.backbone = nn.Sequential(*list(trunk.children())[0:8])   # stem + layer1..4
            if model != 'resnet18':
                self.adjust_dim = nn.Conv2d(1024, 256, kernel_size=1, stride=1, padding=0, bias=False)
                self.bn256 = nn.BatchNorm2d(256)
        elif model == 'facebook':
            # vos_net.py:29-38: torch.hub ResNet-50 (SWSL) de-strided to stride 8 + 2048->1024->256
            trunk = torch.hub.load('facebookresearch/semi-supervised-ImageNet1K-models', 'resnet50_swsl')
            self.backbone = nn.Sequential(*list(trunk.children())[0:8])
            for stage in (6, 7):
                self.backbone[stage][0].conv2.stride = (1, 1)
                self.backbone[stage][0].downsample[0].stride = (1, 1)
            self.adjust_dim = nn.Sequential(nn.Conv2d(2048, 1024, kernel_size=1, bias=False),
                                            nn.Conv2d(1024, 256, kernel_size=1, bias=False))
            self.bn256 = nn.BatchNorm2d(256)
        else:
            raise NotImplementedError

    def forward(self, x):
        x = self.backbone(x)
        if self.model != 'resnet18':
            x = self.bn256(self.adjust_dim(x))
        return x

    def freeze_feature_extraction(self):
        self.backbone.requires_grad_(False)
