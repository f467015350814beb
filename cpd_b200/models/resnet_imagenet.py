"""ImageNet ResNets (the reference's ResNet-50 workload uses torchvision.models.resnet50(),
example/ResNet50/main.py:67).  Own implementation with the same topology and
initialisation; resnet50: 161 parameter tensors, 25 557 032 parameters."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, cin, width, stride=1):
        super().__init__()
        cout = width * self.expansion
        self.conv1 = nn.Conv2d(cin, width, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        self.conv2 = nn.Conv2d(width, width, 3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(width)
        self.conv3 = nn.Conv2d(width, cout, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        y = F.relu(self.bn1(self.conv1(x)), inplace=True)
        y = F.relu(self.bn2(self.conv2(y)), inplace=True)
        y = self.bn3(self.conv3(y))
        return F.relu(y + idt, inplace=True)


class Basic(nn.Module):
    expansion = 1

    def __init__(self, cin, width, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, width, 3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        self.conv2 = nn.Conv2d(width, width, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(width)
        self.downsample = None
        if stride != 1 or cin != width:
            self.downsample = nn.Sequential(nn.Conv2d(cin, width, 1, stride=stride, bias=False), nn.BatchNorm2d(width))

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        y = F.relu(self.bn1(self.conv1(x)), inplace=True)
        y = self.bn2(self.conv2(y))
        return F.relu(y + idt, inplace=True)


class ResNet(nn.Module):
    """ImageNet ResNet-18/34/50/101 with torchvision's parameter layout (the reference uses torchvision.models.resnet50)."""
    def __init__(self, block, layers, num_classes=1000):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
        cin, stages = 64, []
        for i, n in enumerate(layers):
            width = 64 * 2 ** i
            for j in range(n):
                stages.append(block(cin, width, stride=(2 if (i > 0 and j == 0) else 1)))
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
        x = self.maxpool(F.relu(self.bn1(self.conv1(x)), inplace=True))
        x = self.layers(x)
        x = F.adaptive_avg_pool2d(x, 1).flatten(1)
        return self.fc(x)


def resnet18(num_classes=1000):
    return ResNet(Basic, (2, 2, 2, 2), num_classes)


def resnet34(num_classes=1000):
    return ResNet(Basic, (3, 4, 6, 3), num_classes)


def resnet50(num_classes=1000):
    return ResNet(Bottleneck, (3, 4, 6, 3), num_classes)


def resnet101(num_classes=1000):
    return ResNet(Bottleneck, (3, 4, 23, 3), num_classes)
