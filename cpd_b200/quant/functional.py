"""Functional custom-precision API: ``float_quantize``, ``quantizer``, ``quant_gemm``.

Reference: CPDtorch/quant/quant_function.py.  Differences by design: CPU tensors are
supported (the reference raises, :24-30), results are out-of-place by default (the
reference mutates a contiguous input and returns it, quant.cu:14-25), kernels run on
the caller's CUDA stream, sizes are 64-bit, and the GEMM has a tensor-core mode.
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from ..ops import native

__all__ = ["float_quantize", "quantizer", "quant_gemm", "set_default_gemm_mode", "tile_gemm"]


def float_quantize(x: torch.Tensor, exp: int, man: int, rounding: str = "reference", inplace: bool = False,
                   shift: int = 0) -> torch.Tensor:
    """Cast fp32 ``x`` to a float with ``exp`` exponent bits and ``man`` mantissa bits
    (round-to-nearest-even, IEEE-style Inf/NaN, target subnormals) and back to fp32.

    ``rounding="reference"`` is bit-exact with CPD's cast_precision (incl. its quirks,
    SURVEY §2.9.1); ``"ieee"`` is textbook RNE.  ``shift`` multiplies by 2**shift first
    (fused APS scale).  ``(8, 23)`` is the identity except that fp32 denormals flush to 0.
    """
    assert isinstance(x, torch.Tensor), "x is not a single precision Floating Point Tensor"
    if x.dtype != torch.float32:
        raise TypeError(f"float_quantize expects float32, got {x.dtype}")
    if inplace:
        if not x.is_contiguous():
            raise ValueError("inplace float_quantize needs a contiguous tensor")
        return native.quantize_(x, x, exp, man, rounding, shift)
    src = x.contiguous()
    out = torch.empty_like(src)
    return native.quantize_(src, out, exp, man, rounding, shift)


def quantizer(forward_exp: int = 8, forward_man: int = 23, backward_exp: int = 8, backward_man: int = 23,
              rounding: str = "reference"):
    """Autograd function: activations cast to (forward_exp, forward_man) on the way
    forward, gradients cast to (backward_exp, backward_man) on the way back; (8, 23) is
    a pass-through on either side (quant_function.py:33-57)."""

    class Rounding(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            if forward_exp == 8 and forward_man == 23:
                return x
            return float_quantize(x, forward_exp, forward_man, rounding)

        @staticmethod
        def backward(ctx, grad_output):
            if not ctx.needs_input_grad[0]:
                return None
            if backward_exp == 8 and backward_man == 23:
                return grad_output
            return float_quantize(grad_output, backward_exp, backward_man, rounding)

    return Rounding.apply


# ---- GEMM ------------------------------------------------------------------------
_DEFAULT_GEMM_MODE = "exact"


def set_default_gemm_mode(mode: str) -> None:
    """Process-wide default of ``quant_gemm`` / ``Quant_Linear`` / ``Quant_Conv`` when no ``mode`` is given:
    "exact" (reference semantics, the default — also through the CPDtorch facade), "tile" (tensor cores) or "auto"."""
    global _DEFAULT_GEMM_MODE
    if mode not in ("exact", "tile", "auto"):
        raise ValueError(f"unknown gemm mode {mode!r}")
    _DEFAULT_GEMM_MODE = mode


_CPU_MODEL = True            # CPU tensors in tile mode take the CPU model (tests/test_stage_model_cpu.py turns it off to drive
#                              the GPU path's descriptor logic with monkeypatched kernels)
_TILE_MIN_WORK = 1 << 24      # M*N*K above which the tensor-core kernel takes over in mode="auto"


def operand_mode_for(exp: int, man: int, operands: Optional[str] = None) -> int:
    """Tensor-core operand type of the tile mode: fp8 for the fp8-shaped accumulator formats, bf16 up to 7 mantissa
    bits, tf32 above; or what ``operands`` names."""
    if operands:
        return _OPERAND_NAME[operands]
    if (exp, man) == (4, 3):
        return native.OPERAND_E4M3
    if (exp, man) == (5, 2):
        return native.OPERAND_E5M2
    if man <= 7:
        return native.OPERAND_BF16
    return native.OPERAND_TF32


_OPERAND_NAME = {"bf16": native.OPERAND_BF16, "tf32": native.OPERAND_TF32, "e4m3": native.OPERAND_E4M3,
                 "e5m2": native.OPERAND_E5M2}
_OPERAND_FORMAT = {native.OPERAND_BF16: (8, 7), native.OPERAND_TF32: (8, 10), native.OPERAND_E4M3: (4, 3),
                   native.OPERAND_E5M2: (5, 2)}          # (exp, man) of each tensor-core operand type, for the CPU model


def resolve_gemm_mode(mode: Optional[str], is_cuda: bool, work: int) -> str:
    mode = os.environ.get("CPD_B200_GEMM_MODE", mode or _DEFAULT_GEMM_MODE)
    if mode == "auto":
        mode = "tile" if (is_cuda and work >= _TILE_MIN_WORK) else "exact"
    if mode not in ("exact", "tile"):
        raise ValueError(f"unknown gemm mode {mode!r}")
    return mode


def mx_round(x: torch.Tensor, e5m2: bool = False) -> torch.Tensor:
    """What a block-scaled (MXFP8) operand is worth to the tensor core, as fp32: per 32 consecutive K-elements of a row
    one power-of-two scale 2^e, e = ceil(log2(max|x| / max finite code)) (the smallest that avoids saturation - APS per
    block instead of per tensor), values rounded to fp8 after the scaling.  Torch model of mx_quantize_sm100.cu (any
    device); the CUDA kernel additionally emits the codes and the UE8M0 scales in the tcgen05 scale-factor layout."""
    if x.dim() != 2:
        raise ValueError("mx_round: 2-D tensor expected")
    R, K = x.shape
    Kp = (K + 31) // 32 * 32
    xb = torch.nn.functional.pad(x.float(), (0, Kp - K)).view(R, Kp // 32, 32)
    max_code = 57344.0 if e5m2 else 448.0
    amax = xb.abs().amax(-1, keepdim=True)
    mant, ex = torch.frexp(amax / max_code)                        # q = mant * 2^ex, mant in [0.5, 1)
    e = torch.where(mant == 0.5, ex - 1, ex).clamp(-127, 127)       # ceil(log2 q)
    e = torch.where((amax > 0) & torch.isfinite(amax), e, torch.full_like(e, -127))
    scale = torch.ldexp(torch.ones_like(amax), e)
    codes = (xb / scale).clamp(-max_code, max_code).to(torch.float8_e5m2 if e5m2 else torch.float8_e4m3fn).float()
    return (codes * scale).view(R, Kp)[:, :K].contiguous()


def splits_for(M: int, N: int, K: int, k_slice: int) -> int:
    """Split-K factor of the tile mode: skinny outputs with a long reduction (weight gradients: [out_ch, taps] over
    batch x pixels) get enough K ranges to fill the 148 SMs; each split keeps >= 4 K-slices."""
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    slices = max(1, (K + max(k_slice, 1) - 1) // max(k_slice, 1)) if k_slice > 0 else 1
    if tiles >= 96 or slices < 8:
        return 1
    return max(1, min(64, 148 // tiles, slices // 4))


def tile_gemm(a_view: torch.Tensor, bt_view: torch.Tensor, exp: int, man: int, kahan: bool = False, k_slice: int = 128,
              operands: Optional[str] = None, bias: Optional[torch.Tensor] = None, splits: Optional[int] = None) -> torch.Tensor:
    """C = a_view @ bt_view^T for two fp32 2-D VIEWS of any strides, on the tensor cores: each operand goes through ONE
    fused staging launch for BOTH operands (strided gather + cast to the operand type, K-major) and one tensor-core
    launch: the K tail is zero-filled by the TMA descriptor, the bias is added in the GEMM epilogue.  No pad / .to() /
    .contiguous() / transposed copies."""
    om = operand_mode_for(exp, man, operands)
    M, K = a_view.shape
    N = bt_view.shape[0]
    if splits is None:
        splits = splits_for(M, N, K, k_slice)
    if not a_view.is_cuda and _CPU_MODEL:
        # CPU MODEL of the tensor-core semantics (slow; for tests and for reading the numerics without a GPU): operands
        # rounded to the operand format, wide partial sums per K-slice, s = Q(s (+) Q(partial)) with IEEE rounding,
        # split-K = ordered fold of the splits (csrc/cpu/cpu_ops.cpp::gemm_tile_cpu)
        oe, om_ = _OPERAND_FORMAT[om]
        out = native.gemm_tile_cpu(a_view.contiguous(), bt_view.t().contiguous(), exp, man, "ieee", bool(kahan), k_slice, oe, om_, splits)
        return out + bias if bias is not None else out
    a_op, b_op = native.stage(native.spec_matrix(a_view, om), native.spec_matrix(bt_view, om))       # ONE launch, both operands
    return native.gemm_tc(a_op, b_op, exp, man, bool(kahan), k_slice, om, bias, k=K, splits=splits)


def quant_gemm(a: torch.Tensor, b: torch.Tensor, man: int = 23, exp: int = 8, mode: Optional[str] = None,
               kahan: Optional[bool] = None, k_slice: int = 128, operands: Optional[str] = None,
               rounding: str = "reference", bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``a[M,K] @ b[K,N]`` with the accumulator held at (exp, man).  NB: ``man`` before
    ``exp``, as in the reference (quant_function.py:78).

    mode="exact" (default): the reference's semantics — strictly sequential k, every product and
        every (Kahan) intermediate cast; SIMT kernel / C++ on CPU.  Inputs are not cast.
    mode="tile":  B200 throughput path — tcgen05 tensor-core partial sums over ``k_slice``
        elements of K in fp32 (TMEM), then ``s = Q(s (+) Q(partial))`` per slice with IEEE
        rounding (hardware converts).  This mode DOES cast the operands (saturating): fp8 for
        (4,3)/(5,2), bf16 for man<=7, tf32 otherwise — numerics differ from "exact" by design,
        which is why it is opt-in (``mode=``, ``set_default_gemm_mode``, env CPD_B200_GEMM_MODE).
        Operands are staged by one fused gather+cast launch each (no pad / transposed copies).
        On CPU tensors the same semantics run through a (slow) C++ model - operands rounded to the operand format,
        wide partial sums per K-slice, the same casts and split-K fold - so the numerics can be read and tested
        without a GPU (tests/test_tile_mode_cpu.py); block-scaled operands go through ``mx_round``, the torch model of the
        MXFP8 staging kernel.
    mode="auto":  exact on CPU and below 2^24 multiply-adds, tile otherwise.
    ``operands``: "bf16" | "tf32" | "e4m3" | "e5m2" | "mxe4m3" | "mxe5m2" (block-scaled fp8: one
    UE8M0 scale per 32 K-elements, tcgen05 kind::mxf8f6f4.block_scale); default by format.
    ``kahan`` defaults to True for exact (as the reference) and False for tile.
    """
    assert len(a.shape) == 2
    assert len(b.shape) == 2
    assert a.shape[1] == b.shape[0]
    M, K = a.shape
    N = b.shape[1]
    mode = resolve_gemm_mode(mode, a.is_cuda, M * N * K)
    if mode == "exact":
        out = native.gemm_exact(a, b, exp, man, rounding, True if kahan is None else kahan)
        if bias is not None:
            out += bias
        return out
    if not a.is_cuda and operands in ("mxe4m3", "mxe5m2"):
        # CPU model: operands as the tensor core sees them (fp8 codes x their block's power-of-two scale), then the tile model
        e5 = operands == "mxe5m2"
        out = native.gemm_tile_cpu(mx_round(a, e5), mx_round(b.t(), e5).t().contiguous(), exp, man, "ieee",
                                   bool(kahan) if kahan is not None else False, k_slice)
        return out + bias if bias is not None else out
    if operands in ("mxe4m3", "mxe5m2"):       # block-scaled fp8 operands (one power-of-two scale per 32 K-elements)
        e5 = operands == "mxe5m2"
        if K % 16:
            a = torch.nn.functional.pad(a, (0, 16 - K % 16))
            b = torch.nn.functional.pad(b, (0, 0, 0, 16 - K % 16))
        ac, asf = native.mx_quantize(a.contiguous(), e5)
        bc, bsf = native.mx_quantize(b.t().contiguous(), e5)
        return native.gemm_tc_mx(ac, asf, bc, bsf, exp, man, bool(kahan) if kahan is not None else False, k_slice, e5, bias)
    return tile_gemm(a, b.t(), exp, man, bool(kahan) if kahan is not None else False, k_slice, operands, bias)
