"""ResNet-18 for 32x32 inputs (the reference's CIFAR workload: example/ResNet18/models/
resnet18_cifar.py — 3x3 stem, no max-pool, 4 stages x 2 basic blocks, avg-pool 4, fc).
62 parameter tensors, 11 173 962 parameters."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def _conv3x3(cin, cout, stride=1):
    return nn.Conv2d(cin, cout, 3, stride=stride, padding=1, bias=False)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, cin: int, cout: int, stride: int = 1):
        super().__init__()
        self.conv1 = _conv3x3(cin, cout, stride)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = _conv3x3(cout, cout)
        self.bn2 = nn.BatchNorm2d(cout)
        self.proj = None
        if stride != 1 or cin != cout:
            self.proj = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)), inplace=True)
        y = self.bn2(self.conv2(y))
        y = y + (x if self.proj is None else self.proj(x))
        return F.relu(y, inplace=True)


class ResNetCifar(nn.Module):
    """CIFAR ResNet of the reference (resnet18_cifar.py): 3x3 stem without max-pool, 4 stages of BasicBlocks,
    ``avg_pool2d(4)``, linear head; ``forward(x, rank=None)`` keeps the reference's unused ``rank`` argument."""
    def __init__(self, blocks=(2, 2, 2, 2), widths=(64, 128, 256, 512), num_classes: int = 10):
        super().__init__()
        self.stem = nn.Sequential(_conv3x3(3, widths[0]), nn.BatchNorm2d(widths[0]), nn.ReLU(inplace=True))
        stages, cin = [], widths[0]
        for i, (n, w) in enumerate(zip(blocks, widths)):
            for j in range(n):
                stages.append(BasicBlock(cin, w, stride=(2 if (i > 0 and j == 0) else 1)))
                cin = w
        self.stages = nn.Sequential(*stages)
        self.fc = nn.Linear(cin, num_classes)

    def forward(self, x, rank=None):   # `rank` accepted (and ignored) like the reference's forward(x, rank)
        x = self.stages(self.stem(x))
        x = F.avg_pool2d(x, 4).flatten(1)
        return self.fc(x)


def resnet18_cifar(num_classes: int = 10) -> ResNetCifar:
    return ResNetCifar((2, 2, 2, 2), num_classes=num_classes)


def resnet34_cifar(num_classes: int = 10) -> ResNetCifar:
    return ResNetCifar((3, 4, 6, 3), num_classes=num_classes)
