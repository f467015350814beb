"""nn modules over the custom-precision GEMM (reference: CPDtorch/quant/quant_module.py)."""
from __future__ import annotations

import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.nn.init as init
from torch.autograd import Function
from torch.nn.parameter import Parameter

from .functional import (float_quantize, quant_gemm, quantizer, resolve_gemm_mode, operand_mode_for, splits_for,
                         tile_gemm)
from ..ops import native
from . import functional as _functional

__all__ = ["Quantizer", "Quant_Linear", "Quant_Conv", "Quant_LinearFunction", "Quant_ConvTileFunction"]


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
    (quant_module.py:23-52).  Transposes are never materialised: the exact kernel takes strided views, the tile mode
    stages each operand with one gather+cast launch; the bias rides in the tensor-core epilogue and the bias gradient
    is one column-reduce+cast launch."""

    @staticmethod
    def forward(ctx, input, weight, bias=None, exp=8, man=23, mode=None):
        mode = resolve_gemm_mode(mode, input.is_cuda, input.shape[0] * weight.shape[0] * weight.shape[1])
        ctx.save_for_backward(input, weight, bias)
        ctx.exp, ctx.man, ctx.mode = exp, man, mode
        if mode == "tile":
            return tile_gemm(input, weight, exp, man, bias=bias)              # bias added in the epilogue (fp32, un-quantised)
        output = quant_gemm(input, weight.t(), man=man, exp=exp, mode=mode)
        if bias is not None:
            output += bias.unsqueeze(0).expand_as(output)      # fp32, un-quantised, as the reference
        return output

    @staticmethod
    def backward(ctx, grad_output):
        input, weight, bias = ctx.saved_tensors
        grad_input = grad_weight = grad_bias = None
        tile = ctx.mode == "tile"
        if ctx.needs_input_grad[0]:
            grad_input = tile_gemm(grad_output, weight.t(), ctx.exp, ctx.man) if tile else \
                quant_gemm(grad_output, weight, man=ctx.man, exp=ctx.exp, mode=ctx.mode)
        if ctx.needs_input_grad[1]:
            grad_weight = tile_gemm(grad_output.t(), input.t(), ctx.exp, ctx.man) if tile else \
                quant_gemm(grad_output.t(), input, man=ctx.man, exp=ctx.exp, mode=ctx.mode)
        if bias is not None and ctx.needs_input_grad[2]:
            if grad_output.is_cuda and grad_output.stride(1) == 1:
                grad_bias = native.colsum_quantize(grad_output, ctx.exp, ctx.man)
            else:
                grad_bias = float_quantize(grad_output.sum(0), ctx.exp, ctx.man)
        return grad_input, grad_weight, grad_bias, None, None, None


_IMPLICIT = os.environ.get("CPD_B200_CONV_IMPLICIT", "1") != "0"       # implicit-GEMM (im2col-mode TMA) convolutions where covered


def _nhwc_rows(x):
    """[B*H*W, C] view of a channels_last [B, C, H, W] tensor (None when x is laid out otherwise)."""
    B, C, H, W = x.shape
    if x.stride() == (H * W * C, 1, W * C, C) or (C == 1 and x.is_contiguous()):
        return x.as_strided((B * H * W, C), (C, 1), x.storage_offset())
    return None


def _cols_cpu(x, KH, KW, stride, padding, dilation):
    """Explicit im2col matrix [B*OH*OW, (kh, kw, c)] - the CPU model's stand-in for what the GPU path reads through TMA."""
    B, C = x.shape[:2]
    cols = F.unfold(x.contiguous(), (KH, KW), dilation=dilation, padding=padding, stride=stride)        # [B, C*KH*KW, L]
    L = cols.shape[-1]
    return cols.view(B, C, KH * KW, L).permute(0, 3, 2, 1).reshape(B * L, KH * KW * C)


def _conv_tile_cpu(x, weight, bias, exp, man, stride, padding, dilation, k_slice, operands, kahan, dy=None, needs=(True, True, True)):
    """CPU MODEL of the tile-mode convolution (same three GEMMs, K orders, K-slices and split-K factors as the GPU
    path, through ``tile_gemm``'s CPU model).  dy is None: forward, returns y; else returns (dx, dw, db)."""
    B, C, H, W = x.shape
    O, _, KH, KW = weight.shape
    OH = (H + 2 * padding - dilation * (KH - 1) - 1) // stride + 1
    OW = (W + 2 * padding - dilation * (KW - 1) - 1) // stride + 1
    wmat = weight.permute(0, 2, 3, 1).reshape(O, KH * KW * C)                                      # [O, (kh, kw, c)]
    if dy is None:
        y = tile_gemm(_cols_cpu(x, KH, KW, stride, padding, dilation), wmat, exp, man, kahan, k_slice, operands, bias)
        return y.view(B, OH, OW, O).permute(0, 3, 1, 2)
    dx = dw = db = None
    if needs[0]:                                            # transposed convolution: taps of the zero-upsampled dy x flipped filter
        up = dy.new_zeros(B, O, (OH - 1) * stride + 1, (OW - 1) * stride + 1)
        up[:, :, ::stride, ::stride] = dy
        pad_t = dilation * (KH - 1) - padding
        extra_h, extra_w = H - (up.shape[2] + 2 * pad_t - dilation * (KH - 1)), W - (up.shape[3] + 2 * pad_t - dilation * (KW - 1))
        up = F.pad(up, (max(pad_t, 0), max(pad_t, 0) + extra_w, max(pad_t, 0), max(pad_t, 0) + extra_h))
        if pad_t < 0:
            up = up[:, :, -pad_t:up.shape[2] + pad_t, -pad_t:up.shape[3] + pad_t]
        wflip = weight.flip(2, 3).permute(1, 2, 3, 0).reshape(C, KH * KW * O)                      # [C, (kh, kw, o)]
        g = tile_gemm(_cols_cpu(up, KH, KW, 1, 0, dilation), wflip, exp, man, kahan, k_slice, operands)
        dx = g.view(B, H, W, C).permute(0, 3, 1, 2)
    if needs[1]:
        dyt = dy.permute(1, 0, 2, 3).reshape(O, B * OH * OW)                                       # [O, pixels]
        g = tile_gemm(dyt, _cols_cpu(x, KH, KW, stride, padding, dilation).t(), exp, man, kahan, k_slice, operands)
        dw = g.view(O, KH, KW, C).permute(0, 3, 1, 2)
    if bias is not None and needs[2]:
        db = float_quantize(dy.sum((0, 2, 3)), exp, man)
    return dx, dw, db


class Quant_ConvTileFunction(Function):
    """2-D convolution whose three GEMMs run on the tensor cores with the accumulator at (exp, man), WITHOUT the
    reference's glue (quant_module.py:115-139: F.unfold's fp32 column matrix, transposed copies, separate bias add and
    reshapes; autograd's col2im for the data gradient):

      forward   y[(b,oh,ow), o]   = im2col(x)[(b,oh,ow), (kh,kw,c)]  .  W[o, (kh,kw,c)]^T   (+ bias in the epilogue)
      dgrad     dx[(b,h,w), c]    = im2col'(dy)[(b,h,w), (kh,kw,o)]  .  Wflip[c, (kh,kw,o)]^T     (transposed conv:
                                    taps of the zero-upsampled dy, spatially flipped filter — no col2im scatter)
      wgrad     dW[o, (kh,kw,c)]  = dy^T[o, (b,oh,ow)]  .  im2col(x)^T[(kh,kw,c), (b,oh,ow)]^T   (split-K + ordered fold)

    every operand is produced by ONE fused gather/im2col+cast launch from the fp32 tensor in whatever layout it has
    (NCHW or channels_last), both operands of a GEMM in the same launch; the outputs are channels_last views of the
    GEMM results (no permute copies).  Launches: forward 2 (staging + GEMM, bias fused), backward 2 + 3 (split-K fold)
    (+1 bias gradient) - against ~25 torch/extension launches per layer in the reference.

    IMPLICIT GEMM (default where covered: channels_last input, whole 128-byte channel blocks per tap - every ResNet
    layer but the 3-channel stem; ``CPD_B200_CONV_IMPLICIT=0`` turns it off): none of the three im2col matrices is
    written.  The staging launch only casts x / dy; the GEMM kernels read the taps straight from the NHWC tensor with
    im2col-mode TMA loads (forward, stride-1 data gradient) and, for the weight gradient, read dy and the forward's
    bf16 x MN-major in place (``native.gemm_conv`` / ``native.gemm_wgrad``).  Same operand values, K order and
    K-slices as the staged path: bit-identical results (tests/test_gpu_implicit_conv.py).  Launches: forward 2,
    backward 4.  Under-filled deep layers (few output tiles, long K) run split-K with the ordered quantised fold."""

    @staticmethod
    def forward(ctx, x, weight, bias, exp, man, stride, padding, dilation, k_slice, operands, kahan):
        B, C, H, W = x.shape
        O, _, KH, KW = weight.shape
        OH = (H + 2 * padding - dilation * (KH - 1) - 1) // stride + 1
        OW = (W + 2 * padding - dilation * (KW - 1) - 1) // stride + 1
        if not x.is_cuda and _functional._CPU_MODEL:         # CPU model of the same three GEMMs
            ctx.save_for_backward(x, weight)
            ctx.cpu = (exp, man, stride, padding, dilation, k_slice, operands, kahan, bias)
            return _conv_tile_cpu(x, weight, bias, exp, man, stride, padding, dilation, k_slice, operands, kahan)
        ctx.cpu = None
        om = operand_mode_for(exp, man, operands)
        ws = weight.stride()
        wspec = native.spec_gather(weight, O, (KH, KW, C), (ws[2], ws[3], ws[1]), ws[0], om)
        x2 = _nhwc_rows(x)
        implicit = x2 is not None and _IMPLICIT and native.conv_implicit_ok(C, KH, KW, stride, padding, dilation, om)
        # deep layers at small batch leave most SMs idle (32 output tiles, 72 K-blocks each): split-K + ordered fold there too
        fsplits = splits_for(B * OH * OW, O, KH * KW * C, k_slice)
        if implicit:
            # implicit GEMM: the staging launch only CASTS x (2 bytes written per element instead of KH*KW of them)
            a, bt = native.stage(native.spec_matrix(x2, om), wspec)
            y = native.gemm_conv(a.view(B, H, W, C), bt, KH, KW, (OH, OW), stride, padding, dilation, exp, man, kahan, k_slice, om, bias,
                                 splits=fsplits)
        else:
            a, bt = native.stage(native.spec_im2col(x, KH, KW, (OH, OW), stride, padding, dilation, om), wspec)
            y = native.gemm_tc(a, bt, exp, man, kahan, k_slice, om, bias, k=KH * KW * C, splits=fsplits)        # [B*OH*OW, O] == NHWC
        # the weight gradient reads the SAME bf16 copy of x (implicit GEMM again): keep that instead of the fp32 input
        keep_cast = implicit and om == native.OPERAND_BF16 and O % 8 == 0
        ctx.save_for_backward(a.view(B, H, W, C) if keep_cast else x, weight)
        ctx.cfg = (exp, man, stride, padding, dilation, k_slice, om, kahan, bias is not None, (OH, OW), keep_cast, (B, C, H, W))
        return y.view(B, OH, OW, O).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        if ctx.cpu is not None:
            exp, man, stride, padding, dilation, k_slice, operands, kahan, bias = ctx.cpu
            dx, dw, db = _conv_tile_cpu(x, weight, bias, exp, man, stride, padding, dilation, k_slice, operands, kahan, dy=dy,
                                        needs=ctx.needs_input_grad[:3])
            return dx, dw, db, None, None, None, None, None, None, None, None
        exp, man, stride, padding, dilation, k_slice, om, kahan, has_bias, (OH, OW), x_is_cast, (B, C, H, W) = ctx.cfg
        O, _, KH, KW = weight.shape
        dx = dw = db = None
        ws = weight.stride()
        dy2 = _nhwc_rows(dy)
        if x_is_cast and dy2 is None:                       # the implicit weight gradient wants dy as [pixels, O] rows
            dy = dy.contiguous(memory_format=torch.channels_last)
            dy2 = _nhwc_rows(dy)
        dy_cast = None                                      # dy in the operand dtype, [B*OH*OW, O]
        if ctx.needs_input_grad[0]:
            pad_t = dilation * (KH - 1) - padding
            base = (KH - 1) * ws[2] + (KW - 1) * ws[3]
            wspec = native.spec_gather(weight, C, (KH, KW, O), (-ws[2], -ws[3], ws[0]), ws[1], om, base=base)     # flipped filter
            dsplits = splits_for(B * H * W, C, KH * KW * O, k_slice)
            if (dy2 is not None and _IMPLICIT and stride == 1 and pad_t >= 0 and
                    native.conv_implicit_ok(O, KH, KW, 1, pad_t, dilation, om)):
                dy_cast, bt = native.stage(native.spec_matrix(dy2, om), wspec)    # stride-1 data gradient = a convolution of dy
                g = native.gemm_conv(dy_cast.view(B, OH, OW, O), bt, KH, KW, (H, W), 1, pad_t, dilation, exp, man, kahan, k_slice, om,
                                     splits=dsplits)
            else:
                a, bt = native.stage(native.spec_im2col(dy, KH, KW, (H, W), 1, pad_t, dilation, om, in_dil=stride), wspec)   # [(b,h,w), (kh,kw,o)]
                g = native.gemm_tc(a, bt, exp, man, kahan, k_slice, om, None, k=KH * KW * O, splits=dsplits)   # [(b,h,w), c] == NHWC
            dx = g.view(B, H, W, C).permute(0, 3, 1, 2)
        if ctx.needs_input_grad[1]:
            Mred = B * OH * OW
            K = KH * KW * C
            if x_is_cast:
                # implicit weight gradient: dy [pixels, O] and the bf16 x are read MN-major in place
                if dy_cast is None:
                    dy_cast = native.stage(native.spec_matrix(dy2, om))
                g = native.gemm_wgrad(dy_cast, x, KH, KW, (OH, OW), stride, padding, dilation, exp, man, kahan, k_slice,
                                      splits=splits_for(O, K, Mred, k_slice))
            else:
                ds = dy.stride()
                a, bt = native.stage(native.spec_gather(dy, O, (B, OH, OW), (ds[0], ds[2], ds[3]), ds[1], om),        # dy^T  [o, (b,oh,ow)]
                                     native.spec_im2col(x, KH, KW, (OH, OW), stride, padding, dilation, om, taps_are_rows=True))   # [(kh,kw,c), pixels]
                g = native.gemm_tc(a, bt, exp, man, kahan, k_slice, om, None, k=Mred, splits=splits_for(O, K, Mred, k_slice))
            dw = g.view(O, KH, KW, C).permute(0, 3, 1, 2)
        if has_bias and ctx.needs_input_grad[2]:
            if dy.stride(1) == 1 and dy.stride(3) == O and dy.stride(2) == O * OW and dy.stride(0) == O * OW * OH:
                db = native.colsum_quantize(dy.permute(0, 2, 3, 1).reshape(-1, O), exp, man)
            else:
                db = float_quantize(dy.sum((0, 2, 3)), exp, man)
        return dx, dw, db, None, None, None, None, None, None, None, None


class Quant_Linear(nn.Module):
    """``nn.Linear`` whose forward and backward GEMMs keep their accumulator at (exp, man) (quant_module.py:55-85):
    ``mode="exact"`` (default) = the reference's per-MAC casts, ``"tile"`` = tensor cores with a cast per K-slice (casts
    the operands too: opt-in), ``"auto"`` by size; ``None`` follows ``set_default_gemm_mode``."""
    def __init__(self, in_features, out_features, bias=True, exp=8, man=23, mode=None):
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
                 bias=True, exp=8, man=23, mode=None, k_slice=128, operands=None, kahan=False):
        super().__init__()
        if groups != 1:
            raise NotImplementedError("Quant_Conv: groups != 1 is not supported")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = (kernel_size, kernel_size)
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.exp, self.man, self.mode = exp, man, mode
        self.k_slice, self.operands, self.kahan = k_slice, operands, kahan       # tile mode only
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
        mode = resolve_gemm_mode(self.mode, input.is_cuda, batch * oh * ow * self.out_channels * self.in_channels * k * k)
        if mode == "tile":
            return Quant_ConvTileFunction.apply(input, self.weight, self.bias, self.exp, self.man, self.stride, self.padding,
                                                self.dilation, self.k_slice, self.operands, self.kahan)
        cols = F.unfold(input, self.kernel_size, dilation=self.dilation, padding=self.padding, stride=self.stride)
        a = cols.transpose(1, 2).reshape(batch * oh * ow, -1)                  # [B*L, C*k*k]
        y = Quant_LinearFunction.apply(a, self.weight.reshape(self.out_channels, -1), self.bias, self.exp, self.man, mode)
        return y.view(batch, oh * ow, self.out_channels).transpose(1, 2).reshape(batch, self.out_channels, oh, ow)

    def extra_repr(self):
        return "{}, {}, kernel_size={}, stride={}, padding={}, exp={}, man={}".format(
            self.in_channels, self.out_channels, self.kernel_size, self.stride, self.padding, self.exp, self.man)
