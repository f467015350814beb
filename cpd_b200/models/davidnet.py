"""DavidNet (the DAWNBench CIFAR-10 net of the reference's example/DavidNet/davidnet.py +
the graph executor of utils.py:185-292): prep conv, three conv-bn-relu-pool layers of
128/256/512 channels with residual pairs on layers 1 and 3, max-pool 4, bias-free
linear, logits x 0.125.  25 parameter tensors, 6 573 120 parameters.

The reference describes the net as nested dicts executed by a name->(module, inputs)
interpreter that caches every activation.  Here the same wiring is a tiny static
dataflow program (``GraphNet``): nodes are listed once in topological order with
integer input slots, so a forward pass is a flat loop without dict lookups or string
joins, and only the requested outputs are kept.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn as nn


class Scale(nn.Module):
    """Multiplies by a constant (the reference's ``Mul``)."""
    def __init__(self, w: float):
        super().__init__()
        self.w = w

    def forward(self, x):
        return x * self.w


class Add(nn.Module):
    """Element-wise sum of two inputs (residual join)."""
    def forward(self, a, b):
        return a + b


class GraphNet(nn.Module):
    """Static dataflow graph: ``nodes`` = [(name, module, [input names])] in topological
    order; the graph input is called ``"input"``."""

    def __init__(self, nodes: Sequence[Tuple[str, nn.Module, List[str]]], outputs: Sequence[str]):
        super().__init__()
        slot: Dict[str, int] = {"input": 0}
        self.names: List[str] = ["input"]
        self.program: List[Tuple[int, List[int]]] = []
        self.ops = nn.ModuleList()
        for name, mod, ins in nodes:
            if name in slot:
                raise ValueError(f"duplicate node {name}")
            self.ops.append(mod)
            self.program.append((len(self.names), [slot[i] for i in ins]))
            slot[name] = len(self.names)
            self.names.append(name)
        self.out_slots = [slot[o] for o in outputs]
        self.out_names = list(outputs)

    def forward(self, x, rank=None):
        vals: List = [None] * len(self.names)
        vals[0] = x
        for op, (dst, srcs) in zip(self.ops, self.program):
            vals[dst] = op(*[vals[s] for s in srcs])
        outs = [vals[s] for s in self.out_slots]
        return outs[0] if len(outs) == 1 else dict(zip(self.out_names, outs))


def _conv_bn(nodes, name, src, cin, cout, pool=False, fused_bn=False):
    nodes.append((f"{name}_conv", nn.Conv2d(cin, cout, 3, padding=1, bias=False), [src]))
    if fused_bn:                                    # BatchNorm + ReLU in this repo's kernels; the relu node stays as a pass-through
        from ..nn import BatchNormAct2d
        nodes.append((f"{name}_bn", BatchNormAct2d(cout, relu=True), [f"{name}_conv"]))
        nodes.append((f"{name}_relu", nn.Identity(), [f"{name}_bn"]))
    else:
        nodes.append((f"{name}_bn", nn.BatchNorm2d(cout), [f"{name}_conv"]))
        nodes.append((f"{name}_relu", nn.ReLU(inplace=True), [f"{name}_bn"]))
    last = f"{name}_relu"
    if pool:
        nodes.append((f"{name}_pool", nn.MaxPool2d(2), [last]))
        last = f"{name}_pool"
    return last


def _residual(nodes, name, src, c, fused_bn=False):
    a = _conv_bn(nodes, f"{name}_res1", src, c, c, fused_bn=fused_bn)
    b = _conv_bn(nodes, f"{name}_res2", a, c, c, fused_bn=fused_bn)
    nodes.append((f"{name}_add", Add(), [src, b]))
    return f"{name}_add"


def davidnet(channels=(64, 128, 256, 512), num_classes: int = 10, logit_scale: float = 0.125,
             res_layers=(1, 3), fused_bn: bool = True) -> GraphNet:
    nodes: List = []
    cur = _conv_bn(nodes, "prep", "input", 3, channels[0], fused_bn=fused_bn)
    for i in (1, 2, 3):
        cur = _conv_bn(nodes, f"layer{i}", cur, channels[i - 1], channels[i], pool=True, fused_bn=fused_bn)
        if i in res_layers:
            cur = _residual(nodes, f"layer{i}", cur, channels[i], fused_bn=fused_bn)
    nodes.append(("pool", nn.MaxPool2d(4), [cur]))
    nodes.append(("flatten", nn.Flatten(1), ["pool"]))
    nodes.append(("linear", nn.Linear(channels[3], num_classes, bias=False), ["flatten"]))
    nodes.append(("logits", Scale(logit_scale), ["linear"]))
    return GraphNet(nodes, ["logits"])


# ---- dict-described networks (the reference's DavidNet is written this way) ------------------------------
class Identity(nn.Module):
    """Pass-through node."""
    def forward(self, x):
        return x


class Concat(nn.Module):
    """Channel-wise concatenation of its inputs."""
    def forward(self, *xs):
        return torch.cat(xs, 1)


class Correct(nn.Module):
    """Per-sample 0/1 correctness of the arg-max prediction."""
    def forward(self, logits, target):
        return logits.max(dim=1)[1] == target


class RelPath(tuple):
    """Input reference relative to the enclosing dict, e.g. ``rel('in')`` or ``rel('res2', 'relu')``."""


def rel(*parts) -> RelPath:
    return RelPath(parts)


def graph_from_nested(spec: dict, outputs: Sequence[str], sep: str = "_") -> GraphNet:
    """Build a :class:`GraphNet` from a nested-dict description (reference: example/DavidNet/davidnet.py:19-69 with
    the interpreter of utils.py:233-292).  Leaves are ``module`` or ``(module, [inputs])``; a leaf without explicit
    inputs consumes the previous leaf (the first one consumes ``"input"``); inputs are absolute names, tuples of path
    parts, or ``rel(...)`` paths resolved against the leaf's enclosing dict."""
    leaves: List[Tuple[Tuple[str, ...], object]] = []

    def walk(d, pfx):
        for k, v in d.items():
            if isinstance(v, dict):
                walk(v, pfx + (k,))
            else:
                leaves.append((pfx + (k,), v))

    walk(spec, ())
    nodes, prev = [], "input"
    for path, val in leaves:
        name = sep.join(path)
        if isinstance(val, tuple):
            mod, ins = val
            resolved = []
            for i in ins:
                if isinstance(i, RelPath):
                    resolved.append(sep.join(path[:-1] + tuple(i)))
                elif isinstance(i, tuple):
                    resolved.append(sep.join(i))
                else:
                    resolved.append(i)
        else:
            mod, resolved = val, [prev]
        nodes.append((name, mod, resolved))
        prev = name
    return GraphNet(nodes, outputs)


def batch_norm(num_channels, bn_bias_init=None, bn_bias_freeze=False, bn_weight_init=None, bn_weight_freeze=False):
    m = nn.BatchNorm2d(num_channels)
    if bn_bias_init is not None:
        m.bias.data.fill_(bn_bias_init)
    if bn_bias_freeze:
        m.bias.requires_grad = False
    if bn_weight_init is not None:
        m.weight.data.fill_(bn_weight_init)
    if bn_weight_freeze:
        m.weight.requires_grad = False
    return m
