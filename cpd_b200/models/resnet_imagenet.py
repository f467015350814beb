"""ImageNet ResNets (the reference's ResNet-50 workload uses torchvision.models.resnet50(),
example/ResNet50/main.py:67).  Own implementation with the same topology and
initialisation; resnet50: 161 parameter tensors, 25 557 032 parameters."""
from __future__ import annotations

import torch.nn as nn
import torch.nn.functional as F


def _bn(c, fused, relu):
    if fused:
        from ..nn import BatchNormAct2d
        return BatchNormAct2d(c, relu=relu)
    return nn.BatchNorm2d(c)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, cin, width, stride=1, fused_bn=False):
        super().__init__()
        cout = width * self.expansion
        self.fused_bn = fused_bn
        self.conv1 = nn.Conv2d(cin, width, 1, bias=False)
        self.bn1 = _bn(width, fused_bn, True)
        self.conv2 = nn.Conv2d(width, width, 3, stride=stride, padding=1, bias=False)
        self.bn2 = _bn(width, fused_bn, True)
        self.conv3 = nn.Conv2d(width, cout, 1, bias=False)
        self.bn3 = _bn(cout, fused_bn, True)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride, bias=False), _bn(cout, fused_bn, False))

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        if self.fused_bn:                                        # BN+ReLU, BN+ReLU, BN+shortcut+ReLU: 4 launches each
            return self.bn3(self.conv3(self.bn2(self.conv2(self.bn1(self.conv1(x))))), idt)
        y = F.relu(self.bn1(self.conv1(x)), inplace=True)
        y = F.relu(self.bn2(self.conv2(y)), inplace=True)
        y = self.bn3(self.conv3(y))
        return F.relu(y + idt, inplace=True)


class Basic(nn.Module):
    expansion = 1

    def __init__(self, cin, width, stride=1, fused_bn=False):
        super().__init__()
        self.fused_bn = fused_bn
        self.conv1 = nn.Conv2d(cin, width, 3, stride=stride, padding=1, bias=False)
        self.bn1 = _bn(width, fused_bn, True)
        self.conv2 = nn.Conv2d(width, width, 3, padding=1, bias=False)
        self.bn2 = _bn(width, fused_bn, True)
        self.downsample = None
        if stride != 1 or cin != width:
            self.downsample = nn.Sequential(nn.Conv2d(cin, width, 1, stride=stride, bias=False), _bn(width, fused_bn, False))

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        if self.fused_bn:
            return self.bn2(self.conv2(self.bn1(self.conv1(x))), idt)
        y = F.relu(self.bn1(self.conv1(x)), inplace=True)
        y = self.bn2(self.conv2(y))
        return F.relu(y + idt, inplace=True)


class ResNet(nn.Module):
    """ImageNet ResNet-18/34/50/101 with torchvision's parameter layout (the reference uses torchvision.models.resnet50)."""
    def __init__(self, block, layers, num_classes=1000, fused_bn=False):
        super().__init__()
        self.fused_bn = fused_bn
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = _bn(64, fused_bn, True)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
        cin, stages = 64, []
        for i, n in enumerate(layers):
            width = 64 * 2 ** i
            for j in range(n):
                stages.append(block(cin, width, stride=(2 if (i > 0 and j == 0) else 1), fused_bn=fused_bn))
                cin = width * block.expansion
        self.layers = nn.Sequential(*stages)
        self.fc = nn.Linear(cin, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    def forward(self, x, rank=None):
        x = self.maxpool(self.bn1(self.conv1(x)) if self.fused_bn else F.relu(self.bn1(self.conv1(x)), inplace=True))
        x = self.layers(x)
        x = F.adaptive_avg_pool2d(x, 1).flatten(1)
        return self.fc(x)


def resnet18(num_classes=1000, fused_bn=True):
    return ResNet(Basic, (2, 2, 2, 2), num_classes, fused_bn=fused_bn)


def resnet34(num_classes=1000, fused_bn=True):
    return ResNet(Basic, (3, 4, 6, 3), num_classes, fused_bn=fused_bn)


def resnet50(num_classes=1000, fused_bn=True):
    return ResNet(Bottleneck, (3, 4, 6, 3), num_classes, fused_bn=fused_bn)


def resnet101(num_classes=1000, fused_bn=True):
    return ResNet(Bottleneck, (3, 4, 23, 3), num_classes, fused_bn=fused_bn)
