"""BatchNorm2d fused with the residual add and the ReLU that follow it (training mode, channels_last fp32 on CUDA).

The reference's networks (example/ResNet18/models/resnet18_cifar.py:10-45, torchvision ResNet-50 in
example/ResNet50/main.py:67) chain ``conv -> BatchNorm2d -> (+ shortcut) -> ReLU``.  ``BatchNormAct2d`` is a drop-in
``nn.BatchNorm2d`` subclass (same parameters, buffers and state_dict keys, so checkpoints and the "bn"-name weight-decay
rule keep working) whose forward takes the optional residual and applies the ReLU itself: 4 kernel launches per layer for
forward + backward (csrc/cuda/bn_act_sm100.cu) instead of cuDNN BN + add + clamp + threshold + cuDNN BN-backward, and
10-12 passes over the activation tensor instead of 13-16.  Anything the kernels do not cover (eval mode, CPU, NCHW
memory, C % 4 != 0, fp16/bf16) falls back to the stock ops with identical semantics.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

import os

from ..ops import native

# 1-bit ReLU mask for the backward pass instead of re-reading y (CPD_B200_BN_MASK=0 switches it off)
_USE_MASK = os.environ.get("CPD_B200_BN_MASK", "1") == "1"


class _BNActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, res, weight, bias, running_mean, running_var, nbt, momentum, eps, relu, ws):
        B, C, H, W = x.shape
        M = B * H * W
        y = torch.empty_like(x)                                  # keeps the channels_last strides
        stats = torch.empty(4 * C, dtype=torch.float32, device=x.device)
        # 1 bit per element for the backward's ReLU mask (instead of re-reading y there): needs whole 32-element words
        mask = torch.empty(M * C // 32, dtype=torch.int32, device=x.device) if (_USE_MASK and relu and (M * C) % 32 == 0) else None
        with torch.cuda.device(x.device):
            rc = native._C.bn_act_forward_cuda(x.data_ptr(), 0 if res is None else res.data_ptr(), y.data_ptr(), M, C, weight.data_ptr(),
                                               bias.data_ptr(), running_mean.data_ptr(), running_var.data_ptr(),
                                               0 if nbt is None else nbt.data_ptr(), float(momentum), float(eps), bool(relu),
                                               stats.data_ptr(), ws.data_ptr(), 0 if mask is None else mask.data_ptr(),
                                               torch.cuda.current_stream(x.device).cuda_stream)
        if rc != 0:
            raise RuntimeError(f"bn_act_forward: rejected (code {rc})")
        ctx.save_for_backward(x, y if (relu and mask is None) else None, weight, stats, mask)
        ctx.relu, ctx.has_res, ctx.ws = bool(relu), res is not None, ws
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, weight, stats, mask = ctx.saved_tensors
        B, C, H, W = x.shape
        M = B * H * W
        if not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.contiguous(memory_format=torch.channels_last)
        dx = torch.empty_like(x)
        dres = torch.empty_like(x) if (ctx.has_res and ctx.needs_input_grad[1]) else None
        grads = torch.empty(4 * C, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = native._C.bn_act_backward_cuda(dy.data_ptr(), 0 if y is None else y.data_ptr(), 0 if mask is None else mask.data_ptr(),
                                                x.data_ptr(), dx.data_ptr(),
                                                0 if dres is None else dres.data_ptr(), M, C, weight.data_ptr(), stats.data_ptr(),
                                                ctx.relu, grads.data_ptr(), ctx.ws.data_ptr(),
                                                torch.cuda.current_stream(x.device).cuda_stream)
        if rc != 0:
            raise RuntimeError(f"bn_act_backward: rejected (code {rc})")
        return dx, dres, grads[:C], grads[C:2 * C], None, None, None, None, None, None, None


def _channels_last_dense(t: torch.Tensor) -> bool:
    return t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last)


class BatchNormAct2d(nn.BatchNorm2d):
    """``y = relu(batch_norm(x) + residual)`` — ``nn.BatchNorm2d`` with the activation and the shortcut add folded in."""

    def __init__(self, num_features: int, relu: bool = True, **kw):
        super().__init__(num_features, **kw)
        self.relu = bool(relu)
        self._ws: Optional[torch.Tensor] = None

    def _workspace(self, M: int, device) -> torch.Tensor:
        need = int(native._C.bn_workspace_floats(M, self.num_features))
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            self._ws = torch.zeros(need, dtype=torch.float32, device=device)      # counters start at zero; the kernels re-arm them
        return self._ws

    def forward(self, x: torch.Tensor, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
        fused = (self.training and x.is_cuda and x.dtype == torch.float32 and self.affine and self.track_running_stats
                 and self.momentum is not None and self.num_features % 4 == 0 and _channels_last_dense(x)
                 and (residual is None or (residual.dtype == torch.float32 and residual.shape == x.shape
                                           and _channels_last_dense(residual)))
                 and not torch.is_autocast_enabled())
        if fused:
            B, C, H, W = x.shape
            return _BNActFn.apply(x, residual, self.weight, self.bias, self.running_mean, self.running_var, self.num_batches_tracked,
                                  self.momentum, self.eps, self.relu, self._workspace(B * H * W, x.device))
        y = super().forward(x)
        if residual is not None:
            y = y + residual
        return F.relu(y) if self.relu else y

    def extra_repr(self):
        return super().extra_repr() + f", relu={self.relu}"


def fuse_bn_act(bn: nn.BatchNorm2d, relu: bool = True) -> BatchNormAct2d:
    """A ``BatchNormAct2d`` that shares ``bn``'s parameters and buffers (module surgery on an existing network)."""
    m = BatchNormAct2d(bn.num_features, relu=relu, eps=bn.eps, momentum=bn.momentum, affine=bn.affine,
                       track_running_stats=bn.track_running_stats)
    if bn.affine:
        m.weight, m.bias = bn.weight, bn.bias
    if bn.track_running_stats:
        m.running_mean, m.running_var, m.num_batches_tracked = bn.running_mean, bn.running_var, bn.num_batches_tracked
    m.train(bn.training)
    return m
