"""FCN semantic-segmentation network (the reference's fourth workload: FCN-R50-d8 on Cityscapes, trained through an
out-of-tree mmcv / mmsegmentation fork, README.md:132-151).  Own implementation of that architecture: a ResNet whose
last two stages trade their strides for dilations 2 and 4 (output stride 8), an FCN head (3x3 conv -> BN -> ReLU ->
dropout -> 1x1 classifier) on stage 4, the auxiliary head on stage 3, logits upsampled bilinearly to the input size."""
from __future__ import annotations

import torch.nn as nn
import torch.nn.functional as F

from .resnet_imagenet import _bn


class _Block(nn.Module):
    """Bottleneck (expansion 4) or basic (expansion 1) residual block with a dilation."""
    def __init__(self, cin, width, expansion, stride, dilation, fused_bn):
        super().__init__()
        cout = width * expansion
        self.fused_bn = fused_bn
        if expansion == 4:
            convs = [nn.Conv2d(cin, width, 1, bias=False),
                     nn.Conv2d(width, width, 3, stride=stride, padding=dilation, dilation=dilation, bias=False),
                     nn.Conv2d(width, cout, 1, bias=False)]
        else:
            convs = [nn.Conv2d(cin, width, 3, stride=stride, padding=dilation, dilation=dilation, bias=False),
                     nn.Conv2d(width, cout, 3, padding=dilation, dilation=dilation, bias=False)]
        self.convs = nn.ModuleList(convs)
        self.bns = nn.ModuleList([_bn(c.out_channels, fused_bn, True) for c in convs])
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride, bias=False), _bn(cout, fused_bn, False))

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        y = x
        last = len(self.convs) - 1
        for i, (conv, bn) in enumerate(zip(self.convs, self.bns)):
            y = conv(y)
            if self.fused_bn:
                y = bn(y, idt) if i == last else bn(y)               # BN (+ shortcut) + ReLU in one fused module
            else:
                y = bn(y)
                y = F.relu(y + idt if i == last else y, inplace=True)
        return y


class FCNHead(nn.Module):
    def __init__(self, cin, channels, num_classes, dropout=0.1, fused_bn=False):
        super().__init__()
        self.fused_bn = fused_bn
        self.conv = nn.Conv2d(cin, channels, 3, padding=1, bias=False)
        self.bn = _bn(channels, fused_bn, True)
        self.drop = nn.Dropout2d(dropout)
        self.cls = nn.Conv2d(channels, num_classes, 1)

    def forward(self, x):
        y = self.bn(self.conv(x))
        if not self.fused_bn:
            y = F.relu(y, inplace=True)
        return self.cls(self.drop(y))


class FCN(nn.Module):
    """``forward(x)`` returns the main logits at input resolution; in training mode with ``aux=True`` a tuple
    ``(main, aux)`` (loss = CE(main) + 0.4 CE(aux), mmsegmentation's FCN-d8 recipe)."""
    def __init__(self, layers=(3, 4, 6, 3), expansion=4, num_classes=19, aux=True, width=64, fused_bn=False):
        super().__init__()
        self.fused_bn, self.aux = fused_bn, aux
        self.stem = nn.Conv2d(3, width, 7, stride=2, padding=3, bias=False)
        self.stem_bn = _bn(width, fused_bn, True)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
        cin, stages = width, []
        for i, n in enumerate(layers):                              # strides 1, 2, then dilations 2, 4 instead of strides (d8)
            w = width * 2 ** i
            stride, dil = (1, 1) if i == 0 else (2, 1) if i == 1 else (1, 2 ** (i - 1))
            blocks = []
            for j in range(n):
                blocks.append(_Block(cin, w, expansion, stride if j == 0 else 1, dil, fused_bn))
                cin = w * expansion
            stages.append(nn.Sequential(*blocks))
        self.stages = nn.ModuleList(stages)
        self.head = FCNHead(cin, cin // 4, num_classes, fused_bn=fused_bn)
        self.aux_head = FCNHead(cin // 2, cin // 8, num_classes, fused_bn=fused_bn) if aux else None
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    def forward(self, x):
        size = x.shape[-2:]
        y = self.stem_bn(self.stem(x))
        y = self.maxpool(y if self.fused_bn else F.relu(y, inplace=True))
        feats = []
        for st in self.stages:
            y = st(y)
            feats.append(y)
        out = F.interpolate(self.head(feats[3]), size=size, mode="bilinear", align_corners=False)
        if self.training and self.aux_head is not None:
            return out, F.interpolate(self.aux_head(feats[2]), size=size, mode="bilinear", align_corners=False)
        return out


def fcn_r50_d8(num_classes=19, aux=True, fused_bn=False):
    return FCN((3, 4, 6, 3), 4, num_classes, aux, fused_bn=fused_bn)


def fcn_r18_d8(num_classes=19, aux=True, fused_bn=False, width=64):
    return FCN((2, 2, 2, 2), 1, num_classes, aux, width=width, fused_bn=fused_bn)
