"""nn modules over the custom-precision GEMM (reference: CPDtorch/quant/quant_module.py)."""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.nn.init as init
from torch.autograd import Function
from torch.nn.parameter import Parameter

from .functional import float_quantize, quant_gemm, quantizer

__all__ = ["Quantizer", "Quant_Linear", "Quant_Conv", "Quant_LinearFunction"]


class Quantizer(nn.Module):
    """Cast activations on the way forward and gradients on the way back (quant_module.py:13-20)."""

    def __init__(self, forward_exp=8, forward_man=23, backward_exp=8, backward_man=23):
        super().__init__()
        self.fmt = (forward_exp, forward_man, backward_exp, backward_man)
        self.quantize = quantizer(forward_exp, forward_man, backward_exp, backward_man)

    def forward(self, x):
        return self.quantize(x)

    def extra_repr(self):
        return "fwd=(exp={}, man={}), bwd=(exp={}, man={})".format(*self.fmt)


class Quant_LinearFunction(Function):
    """y = x W^T (+ b) with all three GEMMs (fwd, dgrad, wgrad) accumulating at (exp, man)
    (quant_module.py:23-52).  Transposes are passed as strided views — the kernels consume
    them in place, the reference materialises them."""

    @staticmethod
    def forward(ctx, input, weight, bias=None, exp=8, man=23, mode="auto"):
        ctx.save_for_backward(input, weight, bias)
        ctx.exp, ctx.man, ctx.mode = exp, man, mode
        output = quant_gemm(input, weight.t(), man=man, exp=exp, mode=mode)
        if bias is not None:
            output += bias.unsqueeze(0).expand_as(output)      # fp32, un-quantised, as the reference
        return output

    @staticmethod
    def backward(ctx, grad_output):
        input, weight, bias = ctx.saved_tensors
        grad_input = grad_weight = grad_bias = None
        if ctx.needs_input_grad[0]:
            grad_input = quant_gemm(grad_output, weight, man=ctx.man, exp=ctx.exp, mode=ctx.mode)
        if ctx.needs_input_grad[1]:
            grad_weight = quant_gemm(grad_output.t(), input, man=ctx.man, exp=ctx.exp, mode=ctx.mode)
        if bias is not None and ctx.needs_input_grad[2]:
            grad_bias = float_quantize(grad_output.sum(0), ctx.exp, ctx.man)
        return grad_input, grad_weight, grad_bias, None, None, None


class Quant_Linear(nn.Module):
    """``nn.Linear`` whose forward and backward GEMMs keep their accumulator at (exp, man) (quant_module.py:55-85):
    ``mode="exact"`` = the reference's per-MAC casts, ``"tile"`` = tensor cores with a cast per K-slice, ``"auto"`` by size."""
    def __init__(self, in_features, out_features, bias=True, exp=8, man=23, mode="auto"):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.exp, self.man, self.mode = exp, man, mode
        self.weight = Parameter(torch.empty(out_features, in_features))
        if bias:
            self.bias = Parameter(torch.empty(out_features))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = init._calculate_fan_in_and_fan_out(self.weight)
            bound = 1 / math.sqrt(fan_in)
            init.uniform_(self.bias, -bound, bound)

    def forward(self, input):
        lead = input.shape[:-1]
        x2 = input.reshape(-1, self.in_features)
        y = Quant_LinearFunction.apply(x2, self.weight, self.bias, self.exp, self.man, self.mode)
        return y.view(*lead, self.out_features)

    def extra_repr(self):
        return "in_features={}, out_features={}, bias={}, exp={}, man={}".format(
            self.in_features, self.out_features, self.bias is not None, self.exp, self.man)


class Quant_Conv(nn.Module):
    """2-D convolution as im2col + custom-precision GEMM (quant_module.py:88-139).
    Square kernels, int stride/padding/dilation; ``groups`` must be 1 (the reference
    accepts ``dilation``/``groups`` but silently ignores them)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, exp=8, man=23, mode="auto"):
        super().__init__()
        if groups != 1:
            raise NotImplementedError("Quant_Conv: groups != 1 is not supported")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = (kernel_size, kernel_size)
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.exp, self.man, self.mode = exp, man, mode
        self.weight = Parameter(torch.empty(out_channels, in_channels, kernel_size, kernel_size))
        self.bias = Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = init._calculate_fan_in_and_fan_out(self.weight)
            bound = 1 / math.sqrt(fan_in)
            init.uniform_(self.bias, -bound, bound)

    def forward(self, input):
        batch, _c, h, w = input.shape
        k = self.kernel_size[0]
        oh = (h + 2 * self.padding - self.dilation * (k - 1) - 1) // self.stride + 1
        ow = (w + 2 * self.padding - self.dilation * (k - 1) - 1) // self.stride + 1
        cols = F.unfold(input, self.kernel_size, dilation=self.dilation, padding=self.padding, stride=self.stride)
        a = cols.transpose(1, 2).reshape(batch * oh * ow, -1)                  # [B*L, C*k*k]
        y = Quant_LinearFunction.apply(a, self.weight.view(self.out_channels, -1), self.bias, self.exp, self.man, self.mode)
        return y.view(batch, oh * ow, self.out_channels).transpose(1, 2).reshape(batch, self.out_channels, oh, ow)

    def extra_repr(self):
        return "{}, {}, kernel_size={}, stride={}, padding={}, exp={}, man={}".format(
            self.in_channels, self.out_channels, self.kernel_size, self.stride, self.padding, self.exp, self.man)
