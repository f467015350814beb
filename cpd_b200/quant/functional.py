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

__all__ = ["float_quantize", "quantizer", "quant_gemm"]


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
_TILE_MIN_WORK = 1 << 24      # M*N*K above which the tensor-core kernel takes over in mode="auto"


def _operand_mode(exp: int, man: int) -> int:
    if (exp, man) == (4, 3):
        return native.OPERAND_E4M3
    if (exp, man) == (5, 2):
        return native.OPERAND_E5M2
    if man <= 7:
        return native.OPERAND_BF16
    return native.OPERAND_TF32


_OPERAND_DTYPE = {native.OPERAND_BF16: torch.bfloat16, native.OPERAND_TF32: torch.float32,
                  native.OPERAND_E4M3: torch.float8_e4m3fn, native.OPERAND_E5M2: torch.float8_e5m2}
_OPERAND_NAME = {"bf16": native.OPERAND_BF16, "tf32": native.OPERAND_TF32, "e4m3": native.OPERAND_E4M3,
                 "e5m2": native.OPERAND_E5M2}


def quant_gemm(a: torch.Tensor, b: torch.Tensor, man: int = 23, exp: int = 8, mode: str = "auto",
               kahan: Optional[bool] = None, k_slice: int = 128, operands: Optional[str] = None,
               rounding: str = "reference", bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``a[M,K] @ b[K,N]`` with the accumulator held at (exp, man).  NB: ``man`` before
    ``exp``, as in the reference (quant_function.py:78).

    mode="exact": the reference's semantics — strictly sequential k, every product and
        every (Kahan) intermediate cast; SIMT kernel / C++ on CPU.  Inputs are not cast.
    mode="tile":  B200 throughput path — tcgen05 tensor-core partial sums over ``k_slice``
        elements of K in fp32 (TMEM), then ``s = Q(s (+) Q(partial))`` per slice with IEEE
        rounding (hardware converts); operands are staged as fp8 for (4,3)/(5,2), bf16 for
        man<=7, tf32 otherwise.  Large problems in the 16-bit-representable formats run on
        CTA pairs (thread-block clusters, cta_group::2) with packed 16-bit accumulators.
    mode="auto":  exact on CPU and for small problems, tile otherwise.
    ``operands``: "bf16" | "tf32" | "e4m3" | "e5m2" | "mxe4m3" | "mxe5m2" (block-scaled fp8: one
    UE8M0 scale per 32 K-elements, tcgen05 kind::mxf8f6f4.block_scale); default by format.
    ``kahan`` defaults to True for exact (as the reference) and False for tile.
    """
    assert len(a.shape) == 2
    assert len(b.shape) == 2
    assert a.shape[1] == b.shape[0]
    mode = os.environ.get("CPD_B200_GEMM_MODE", mode)
    M, K = a.shape
    N = b.shape[1]
    if mode == "auto":
        mode = "tile" if (a.is_cuda and M * N * K >= _TILE_MIN_WORK) else "exact"
    if mode == "exact":
        out = native.gemm_exact(a, b, exp, man, rounding, True if kahan is None else kahan)
        if bias is not None:
            out += bias
        return out
    if mode != "tile":
        raise ValueError(f"unknown gemm mode {mode!r}")
    if not a.is_cuda:
        raise NotImplementedError("mode='tile' needs a CUDA device (use ops.native.gemm_tile_cpu for the CPU model)")
    if K % 16:      # TMA needs 16-byte aligned rows: zero-pad K (e.g. the 3x3x3 = 27 wide im2col of a first conv layer)
        padk = 16 - K % 16
        a = torch.nn.functional.pad(a, (0, padk))
        b = torch.nn.functional.pad(b, (0, 0, 0, padk))
    if operands in ("mxe4m3", "mxe5m2"):       # block-scaled fp8 operands (one power-of-two scale per 32 K-elements)
        e5 = operands == "mxe5m2"
        ac, asf = native.mx_quantize(a.contiguous(), e5)
        bc, bsf = native.mx_quantize(b.t().contiguous(), e5)
        return native.gemm_tc_mx(ac, asf, bc, bsf, exp, man, bool(kahan) if kahan is not None else False, k_slice, e5, bias)
    om = _OPERAND_NAME[operands] if operands else _operand_mode(exp, man)
    dt = _OPERAND_DTYPE[om]
    a_op = a.to(dt).contiguous()
    bt_op = b.t().to(dt).contiguous()           # [N, K], K-major: no copy if b is itself a transposed view
    return native.gemm_tc(a_op, bt_op, exp, man, bool(kahan) if kahan is not None else False, k_slice, om, bias)
