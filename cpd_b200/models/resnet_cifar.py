"""ResNet-18 for 32x32 inputs (the reference's CIFAR workload: example/ResNet18/models/
resnet18_cifar.py — 3x3 stem, no max-pool, 4 stages x 2 basic blocks, avg-pool 4, fc).
62 parameter tensors, 11 173 962 parameters.  ``fused_bn`` (default): the BatchNorm layers are ``BatchNormAct2d`` -
the same parameters, buffers and arithmetic with the ReLU and the shortcut add folded into this repo's kernels on
channels_last fp32 CUDA tensors (stock ops everywhere else)."""
from __future__ import annotations

import torch.nn as nn
import torch.nn.functional as F


def _plain_conv(cin, cout, k, stride=1, padding=0):
    return nn.Conv2d(cin, cout, k, stride=stride, padding=padding, bias=False)


def _bn(c, fused, relu):
    """nn.BatchNorm2d, or (fused_bn) the same module with the ReLU / shortcut add folded into its kernels."""
    if fused:
        from ..nn import BatchNormAct2d
        return BatchNormAct2d(c, relu=relu)
    return nn.BatchNorm2d(c)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, cin: int, cout: int, stride: int = 1, conv=_plain_conv, fused_bn: bool = False):
        super().__init__()
        self.fused_bn = fused_bn
        self.conv1 = conv(cin, cout, 3, stride=stride, padding=1)
        self.bn1 = _bn(cout, fused_bn, relu=True)
        self.conv2 = conv(cout, cout, 3, stride=1, padding=1)
        self.bn2 = _bn(cout, fused_bn, relu=True)
        self.proj = None
        if stride != 1 or cin != cout:
            self.proj = nn.Sequential(conv(cin, cout, 1, stride=stride, padding=0), _bn(cout, fused_bn, relu=False))

    def forward(self, x):
        shortcut = x if self.proj is None else self.proj(x)
        if self.fused_bn:
            y = self.bn1(self.conv1(x))                         # BN + ReLU in one pass pair
            return self.bn2(self.conv2(y), shortcut)            # BN + shortcut add + ReLU
        y = F.relu(self.bn1(self.conv1(x)), inplace=True)
        y = self.bn2(self.conv2(y))
        y = y + shortcut
        return F.relu(y, inplace=True)


class ResNetCifar(nn.Module):
    """CIFAR ResNet of the reference (resnet18_cifar.py): 3x3 stem without max-pool, 4 stages of BasicBlocks,
    ``avg_pool2d(4)``, linear head; ``forward(x, rank=None)`` keeps the reference's unused ``rank`` argument."""
    def __init__(self, blocks=(2, 2, 2, 2), widths=(64, 128, 256, 512), num_classes: int = 10, conv=_plain_conv,
                 fused_bn: bool = False):
        super().__init__()
        # same parameter / buffer names either way: the fused stem keeps an (identity) third slot
        self.stem = nn.Sequential(conv(3, widths[0], 3, stride=1, padding=1), _bn(widths[0], fused_bn, relu=True),
                                  nn.Identity() if fused_bn else nn.ReLU(inplace=True))
        stages, cin = [], widths[0]
        for i, (n, w) in enumerate(zip(blocks, widths)):
            for j in range(n):
                stages.append(BasicBlock(cin, w, stride=(2 if (i > 0 and j == 0) else 1), conv=conv, fused_bn=fused_bn))
                cin = w
        self.stages = nn.Sequential(*stages)
        self.fc = nn.Linear(cin, num_classes)

    def forward(self, x, rank=None):   # `rank` accepted (and ignored) like the reference's forward(x, rank)
        x = self.stages(self.stem(x))
        x = F.avg_pool2d(x, 4).flatten(1)
        return self.fc(x)


def resnet18_cifar(num_classes: int = 10, fused_bn: bool = True) -> ResNetCifar:
    return ResNetCifar((2, 2, 2, 2), num_classes=num_classes, fused_bn=fused_bn)


def resnet34_cifar(num_classes: int = 10, fused_bn: bool = True) -> ResNetCifar:
    return ResNetCifar((3, 4, 6, 3), num_classes=num_classes, fused_bn=fused_bn)


def resnet18_cifar_quant(num_classes: int = 10, exp: int = 5, man: int = 10, mode: str = "tile", operands=None, k_slice: int = 128,
                         kahan: bool = False, fused_bn: bool = True) -> ResNetCifar:
    """The same network with EVERY convolution a ``Quant_Conv`` (CPDtorch/quant/quant_module.py:88-139): forward,
    data-gradient and weight-gradient GEMMs keep their accumulator at (exp, man).  ``mode="tile"`` runs them on the
    tcgen05 kernels with fused im2col/cast staging, ``"exact"`` is the reference's per-MAC semantics.  The 10-class
    head stays ``nn.Linear``: the reference's GEMM is unreliable for outputs with fewer than 16 rows (its remainder
    branch), so the comparison arm cannot quantise it either."""
    from ..quant import Quant_Conv

    def conv(cin, cout, k, stride=1, padding=0):
        return Quant_Conv(cin, cout, k, stride=stride, padding=padding, bias=False, exp=exp, man=man, mode=mode,
                          operands=operands, k_slice=k_slice, kahan=kahan)
    return ResNetCifar((2, 2, 2, 2), num_classes=num_classes, conv=conv, fused_bn=fused_bn)
