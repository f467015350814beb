"""Validated python entry points of the native ops (CPU C++ and sm_100a CUDA).

Every function checks dtype / device / contiguity, picks the current CUDA stream of
the tensor's device and forwards raw pointers to ``cpd_b200._C`` (which is torch-free).
CUDA tensors always run the CUDA kernels; there is no eager-PyTorch fallback.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from .._ext import load

_C = load()

ROUND_REF = 0
ROUND_IEEE = 1
_ROUNDING = {"reference": ROUND_REF, "ref": ROUND_REF, "ieee": ROUND_IEEE, 0: 0, 1: 1}

CHUNK_ELEMS = 4096          # elements per scheduling chunk (16 per thread x 256 threads)
ALIGN_ELEMS = 16            # tensors start on 64-byte boundaries inside an arena
CHUNK_DTYPE = np.dtype([("start", "<i8"), ("len", "<i4"), ("tensor", "<i4")])
assert CHUNK_DTYPE.itemsize == _C.chunk_struct_size()

ALGO_AUTO, ALGO_ONESHOT, ALGO_TWOSHOT, ALGO_NVLS, ALGO_ONESHOT_RAW, ALGO_NVLS_SUM, ALGO_NVLS_FAST = 0, 1, 2, 3, 4, 5, 6


def rounding_code(rounding) -> int:
    try:
        return _ROUNDING[rounding]
    except KeyError:
        raise ValueError(f"rounding must be 'reference' or 'ieee', got {rounding!r}") from None


def _check_fmt(exp: int, man: int) -> None:
    if not (1 <= int(exp) <= 8):
        raise ValueError(f"exp must be in [1, 8], got {exp}")
    if not (0 <= int(man) <= 23):
        raise ValueError(f"man must be in [0, 23], got {man}")


def _f32(t: torch.Tensor, name: str) -> None:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} is not a torch.Tensor")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


def wire_bits(exp: int, man: int) -> int:
    return _C.wire_bits(int(exp), int(man))


# ------------------------------------------------------------------ elementwise --
def quantize_(inp: torch.Tensor, out: torch.Tensor, exp: int, man: int, rounding=ROUND_REF, shift: int = 0) -> torch.Tensor:
    """out = Q_(exp,man)(inp * 2**shift); ``out`` may alias ``inp``.  Both contiguous fp32."""
    _f32(inp, "input"); _f32(out, "out"); _check_fmt(exp, man)
    if not (inp.is_contiguous() and out.is_contiguous()):
        raise ValueError("quantize_: tensors must be contiguous")
    if inp.numel() != out.numel() or inp.device != out.device:
        raise ValueError("quantize_: input/out mismatch")
    if not (-126 <= shift <= 127):
        raise ValueError("shift out of range")
    n = inp.numel()
    if n == 0:
        return out
    r = rounding_code(rounding)
    if inp.is_cuda:
        with torch.cuda.device(inp.device):
            _C.quantize_cuda(inp.data_ptr(), out.data_ptr(), n, int(exp), int(man), r, int(shift), _stream(inp))
    else:
        _C.quantize_cpu(inp.data_ptr(), out.data_ptr(), n, int(exp), int(man), r, int(shift))
    return out


def amax_bits(inp: torch.Tensor, out_u32: Optional[torch.Tensor] = None) -> torch.Tensor:
    """max |x| as fp32 bit pattern (int32 tensor of 1 element), CUDA only."""
    _f32(inp, "input")
    if not inp.is_cuda or not inp.is_contiguous():
        raise ValueError("amax_bits: contiguous CUDA tensor expected")
    if out_u32 is None:
        out_u32 = torch.zeros(1, dtype=torch.int32, device=inp.device)
    with torch.cuda.device(inp.device):
        _C.amax_cuda(inp.data_ptr(), inp.numel(), out_u32.data_ptr(), _stream(inp))
    return out_u32


# ----------------------------------------------------------------------- GEMM --
def gemm_exact(a: torch.Tensor, b: torch.Tensor, exp: int, man: int, rounding=ROUND_REF, kahan: bool = True,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """C = A @ B with the reference's per-MAC quantised (Kahan) sequential-k accumulation.
    A and B may have arbitrary 2-D strides (transposed views are consumed in place)."""
    _f32(a, "a"); _f32(b, "b"); _check_fmt(exp, man)
    if a.dim() != 2 or b.dim() != 2 or a.shape[1] != b.shape[0]:
        raise ValueError(f"gemm: bad shapes {tuple(a.shape)} x {tuple(b.shape)}")
    if a.device != b.device:
        raise ValueError("gemm: device mismatch")
    M, K = a.shape
    N = b.shape[1]
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=a.device)
    else:
        _f32(out, "out")
        if tuple(out.shape) != (M, N) or not out.is_contiguous() or out.device != a.device:
            raise ValueError("gemm: bad out tensor")
    if M == 0 or N == 0:
        return out
    r = rounding_code(rounding)
    args = (a.data_ptr(), a.stride(0), a.stride(1), b.data_ptr(), b.stride(0), b.stride(1), out.data_ptr(), out.stride(0),
            M, N, K, int(exp), int(man), r, bool(kahan))
    if a.is_cuda:
        with torch.cuda.device(a.device):
            _C.gemm_exact_cuda(*args, _stream(a))
    else:
        _C.gemm_exact_cpu(*args)
    return out


def gemm_tile_cpu(a, b, exp, man, rounding=ROUND_REF, kahan=False, k_slice=128, op_exp=8, op_man=23, splits=1):
    """CPU model of the tensor-core kernel's semantics (per-K-slice requantisation; split-K = ordered fold of the splits)."""
    _f32(a, "a"); _f32(b, "b"); _check_fmt(exp, man)
    M, K = a.shape
    N = b.shape[1]
    out = torch.empty(M, N, dtype=torch.float32)
    _C.gemm_tile_cpu(a.data_ptr(), a.stride(0), a.stride(1), b.data_ptr(), b.stride(0), b.stride(1), out.data_ptr(),
                     out.stride(0), M, N, K, int(exp), int(man), rounding_code(rounding), bool(kahan), int(k_slice),
                     int(op_exp), int(op_man), int(splits))
    return out


OPERAND_BF16, OPERAND_TF32, OPERAND_E4M3, OPERAND_E5M2 = 0, 1, 2, 3


GEMM_AUTO, GEMM_1CTA, GEMM_2CTA = 0, 1, 2


_OPERAND_TORCH = {OPERAND_BF16: torch.bfloat16, OPERAND_TF32: torch.float32, OPERAND_E4M3: torch.float8_e4m3fn,
                  OPERAND_E5M2: torch.float8_e5m2}


def gemm_tc(a: torch.Tensor, bt: torch.Tensor, exp: int, man: int, kahan: bool = False, k_slice: int = 128,
            operand_mode: int = OPERAND_BF16, bias: Optional[torch.Tensor] = None,
            out: Optional[torch.Tensor] = None, variant: int = GEMM_AUTO, k: Optional[int] = None,
            splits: int = 1) -> torch.Tensor:
    """tcgen05 GEMM: C[M,N] = A[M,K] @ Bt[N,K]^T, accumulator re-quantised to (exp,man)
    (IEEE rounding) after every K-slice.  Operands must already be in the operand dtype (bf16 /
    fp32 for tf32 / float8), K-major with unit inner stride; rows may be padded (``k`` = the true K, the TMA
    descriptor zero-fills the tail).  ``variant``: auto | 128x128 single-CTA tiles | 256x256 CTA-pair
    (cluster, cta_group::2) tiles with packed 16-bit accumulators.  ``splits`` > 1: split-K with an ordered
    quantised fold of the partial matrices (skinny outputs with a long reduction, e.g. weight gradients)."""
    _check_fmt(exp, man)
    if not (a.is_cuda and bt.is_cuda and a.dim() == 2 and bt.dim() == 2 and a.stride(1) == 1 and bt.stride(1) == 1):
        raise ValueError("gemm_tc: 2-D CUDA operands with unit inner stride expected")
    want = _OPERAND_TORCH[operand_mode]
    if a.dtype != want or bt.dtype != want:
        raise TypeError(f"gemm_tc: operand_mode {operand_mode} needs {want}, got {a.dtype}/{bt.dtype}")
    M, N = a.shape[0], bt.shape[0]
    K = int(k) if k is not None else a.shape[1]
    if K > a.shape[1] or K > bt.shape[1] or (k is None and bt.shape[1] != K):
        raise ValueError("gemm_tc: K mismatch")
    lda, ldb = (a.stride(0) if M > 1 else a.shape[1]), (bt.stride(0) if N > 1 else bt.shape[1])
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=a.device)
    ws = None
    splits = max(1, int(splits))
    if splits > 1:
        ws = torch.empty(splits, M, N, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        rc = _C.gemm_tc_cuda(a.data_ptr(), bt.data_ptr(), out.data_ptr(), M, N, K, int(exp), int(man), bool(kahan),
                             int(k_slice), int(operand_mode), _ptr(bias), int(variant), _stream(a), int(lda), int(ldb),
                             splits, _ptr(ws))
    if rc != 0:
        raise RuntimeError(f"gemm_tc: kernel rejected the problem (code {rc})")
    return out


def conv_implicit_ok(C: int, kh: int, kw: int, stride: int, pad: int, dil: int, operand_mode: int) -> bool:
    """Does the implicit-GEMM (im2col-mode TMA) path cover this convolution?  Whole 128-byte channel blocks per tap and
    the descriptor's corner / offset ranges; everything else goes through the explicit staging kernel."""
    cpb = 128 // _OPERAND_TORCH[operand_mode].itemsize
    return (C > 0 and C % cpb == 0 and 1 <= stride <= 8 and dil >= 1 and 0 <= pad <= 127 and (kh - 1) * dil <= 255 and
            (kw - 1) * dil <= 255 and pad - (kh - 1) * dil >= -128 and pad - (kw - 1) * dil >= -128)


def gemm_conv(x_nhwc: torch.Tensor, bt: torch.Tensor, kh: int, kw: int, out_hw, stride: int, pad: int, dil: int, exp: int, man: int,
              kahan: bool = False, k_slice: int = 128, operand_mode: int = OPERAND_BF16, bias: Optional[torch.Tensor] = None,
              variant: int = GEMM_AUTO, splits: int = 1) -> torch.Tensor:
    """Implicit-GEMM convolution on tcgen05: y[(b,oh,ow), o] = sum_(kh,kw,c) x[b, oh*s-p+kh*d, ow*s-p+kw*d, c] * bt[o, (kh,kw,c)].
    ``x_nhwc`` is the dense [B, H, W, C] activation tensor ALREADY in the operand dtype; the im2col matrix is never
    written: the kernel's TMA producer reads the taps straight from it (im2col-mode tensor map, zero-filled halo).
    ``bt`` [O, >= KH*KW*C] K-major filter matrix.  Accumulator re-quantised to (exp, man) every ``k_slice``."""
    _check_fmt(exp, man)
    want = _OPERAND_TORCH[operand_mode]
    if not (x_nhwc.is_cuda and x_nhwc.dim() == 4 and x_nhwc.is_contiguous() and x_nhwc.dtype == want):
        raise ValueError(f"gemm_conv: dense [B, H, W, C] {want} CUDA tensor expected")
    B, H, W, C = x_nhwc.shape
    OH, OW = int(out_hw[0]), int(out_hw[1])
    K = kh * kw * C
    if not (bt.is_cuda and bt.dim() == 2 and bt.stride(1) == 1 and bt.dtype == want and bt.shape[1] >= K):
        raise ValueError("gemm_conv: K-major filter matrix [O, KH*KW*C] in the operand dtype expected")
    if not conv_implicit_ok(C, kh, kw, stride, pad, dil, operand_mode):
        raise ValueError("gemm_conv: convolution not covered by the implicit-GEMM path (see conv_implicit_ok)")
    O = bt.shape[0]
    out = torch.empty(B * OH * OW, O, dtype=torch.float32, device=x_nhwc.device)
    splits = max(1, int(splits))
    ws = torch.empty(splits, B * OH * OW, O, dtype=torch.float32, device=x_nhwc.device) if splits > 1 else None
    with torch.cuda.device(x_nhwc.device):
        rc = _C.gemm_conv_cuda(x_nhwc.data_ptr(), bt.data_ptr(), out.data_ptr(), O, [B, H, W, C, kh, kw, OH, OW, stride, pad, dil],
                               int(exp), int(man), bool(kahan), int(k_slice), int(operand_mode), _ptr(bias), int(variant),
                               _stream(x_nhwc), int(bt.stride(0) if O > 1 else bt.shape[1]), splits, _ptr(ws))
    if rc != 0:
        raise RuntimeError(f"gemm_conv: kernel rejected the problem (code {rc})")
    return out


def gemm_wgrad(dy_rows: torch.Tensor, x_nhwc: torch.Tensor, kh: int, kw: int, out_hw, stride: int, pad: int, dil: int, exp: int,
               man: int, kahan: bool = False, k_slice: int = 128, splits: int = 1, _lbo: int = 0, _sbo: int = 0) -> torch.Tensor:
    """Weight gradient of a convolution as an implicit GEMM: dW[o, (kh,kw,c)] = sum_pixels dy[pixel, o] * x-tap[pixel, (kh,kw,c)].
    ``dy_rows`` [B*OH*OW, O] and ``x_nhwc`` [B, H, W, C] dense bf16.  Both operands are read MN-major straight from those
    tensors (tiled TMA for dy, im2col-mode TMA for x): no transposed copy of dy, no (transposed) im2col matrix.  The long
    reduction over pixels runs split-K with the ordered quantised fold of ``gemm_tc``."""
    _check_fmt(exp, man)
    if not (x_nhwc.is_cuda and x_nhwc.dim() == 4 and x_nhwc.is_contiguous() and x_nhwc.dtype == torch.bfloat16):
        raise ValueError("gemm_wgrad: dense [B, H, W, C] bf16 CUDA tensor expected")
    B, H, W, C = x_nhwc.shape
    OH, OW = int(out_hw[0]), int(out_hw[1])
    if not (dy_rows.is_cuda and dy_rows.dim() == 2 and dy_rows.is_contiguous() and dy_rows.dtype == torch.bfloat16 and
            dy_rows.shape[0] == B * OH * OW):
        raise ValueError("gemm_wgrad: dense [B*OH*OW, O] bf16 gradient expected")
    O = dy_rows.shape[1]
    if O % 8 or not conv_implicit_ok(C, kh, kw, stride, pad, dil, OPERAND_BF16):
        raise ValueError("gemm_wgrad: convolution not covered by the implicit-GEMM path")
    N = kh * kw * C
    out = torch.empty(O, N, dtype=torch.float32, device=x_nhwc.device)
    splits = max(1, int(splits))
    ws = torch.empty(splits, O, N, dtype=torch.float32, device=x_nhwc.device) if splits > 1 else None
    with torch.cuda.device(x_nhwc.device):
        rc = _C.gemm_wgrad_cuda(dy_rows.data_ptr(), x_nhwc.data_ptr(), out.data_ptr(), O, [B, H, W, C, kh, kw, OH, OW, stride, pad, dil],
                                int(exp), int(man), bool(kahan), int(k_slice), _stream(x_nhwc), splits, _ptr(ws), int(_lbo), int(_sbo))
    if rc != 0:
        raise RuntimeError(f"gemm_wgrad: kernel rejected the problem (code {rc})")
    return out


# ---- operand staging: gather / im2col + cast in one kernel (cuda/operand_stage_sm100.cu) -----------------------------
def _staged_empty(R: int, K: int, operand_mode: int, device) -> torch.Tensor:
    Kp = (K + 15) // 16 * 16           # 16-byte rows for every operand type (TMA global stride)
    return torch.empty(R, Kp, dtype=_OPERAND_TORCH[operand_mode], device=device)


class StageSpec:
    """One operand of a tile-mode GEMM: an address-generator descriptor (gather or im2col) over an fp32 CUDA tensor,
    its logical shape [R, K] and the output buffer.  ``stage(spec_a, spec_b)`` fills both in ONE launch."""

    def __init__(self, src: torch.Tensor, desc: bytes, R: int, K: int, operand_mode: int):
        self.src, self.desc, self.R, self.K, self.operand_mode = src, desc, int(R), int(K), int(operand_mode)
        self.out = _staged_empty(self.R, self.K, operand_mode, src.device)


def spec_gather(src: torch.Tensor, R: int, sizes, strides, s_row: int, operand_mode: int, base: int = 0,
                rows_fast: Optional[bool] = None) -> StageSpec:
    """out[r, k] = src.flat[base + r * s_row + k1 * s1 + k2 * s2 + k3 * s3], k = (k1, k2, k3) row-major over ``sizes``,
    cast to the operand dtype.  Strides are in elements of ``src``'s storage view and may be negative."""
    _f32(src, "src")
    if not src.is_cuda:
        raise ValueError("stage: CUDA tensor expected")
    n1, n2, n3 = [int(v) for v in sizes]
    s1, s2, s3 = [int(v) for v in strides]
    if rows_fast is None:
        rows_fast = abs(int(s_row)) == 1 and abs(s3) != 1
    desc = _C.stage_gather_desc(src.data_ptr(), int(base), int(bool(rows_fast)), int(s_row), s1, s2, s3, n2, n3)
    return StageSpec(src, desc, R, n1 * n2 * n3, operand_mode)


def spec_matrix(x: torch.Tensor, operand_mode: int) -> StageSpec:
    """Any 2-D fp32 view (transposed, sliced, ...) as a K-major operand: no materialised transpose, no pad."""
    if x.dim() != 2:
        raise ValueError("stage: 2-D tensor expected")
    return spec_gather(x, x.shape[0], (x.shape[1], 1, 1), (x.stride(1), 0, 0), x.stride(0), operand_mode,
                       rows_fast=(abs(x.stride(0)) == 1 and abs(x.stride(1)) != 1))


def spec_im2col(x: torch.Tensor, kh: int, kw: int, out_hw, stride: int, pad: int, dil: int, operand_mode: int,
                in_dil: int = 1, taps_are_rows: bool = False) -> StageSpec:
    """im2col of x[B, C, H, W] (any strides: NCHW or channels_last) fused with the operand cast:
    rows = output pixels (b, oh, ow), k = taps (kh, kw, c)  [taps_are_rows: the transposed matrix].  ``in_dil`` > 1
    treats x as zero-upsampled by that factor (data-gradient of a strided convolution)."""
    _f32(x, "x")
    if not x.is_cuda or x.dim() != 4:
        raise ValueError("stage: 4-D CUDA tensor expected")
    B, C, H, W = x.shape
    OH, OW = int(out_hw[0]), int(out_hw[1])
    pixels, taps = B * OH * OW, kh * kw * C
    R, K = (taps, pixels) if taps_are_rows else (pixels, taps)
    sb, sc, sh, sw = x.stride()
    # which logical dimension is contiguous in the source?  channels_last: c = the fastest TAP index; NCHW: iw ~ the
    # fastest PIXEL index.  The kernel reads that dimension across the threads of a warp.
    chan_fast = sc == 1 and C > 1
    rows_fast = chan_fast == bool(taps_are_rows)
    desc = _C.stage_im2col_desc(x.data_ptr(), 0, int(rows_fast), int(bool(taps_are_rows)), C, H, W, OH, OW, int(kh), int(kw),
                                int(stride), int(pad), int(dil), int(in_dil), sb, sh, sw, sc)
    return StageSpec(x, desc, R, K, operand_mode)


def stage(*specs: StageSpec):
    """Fill the operand buffers of one or two specs with ONE kernel launch; returns the staged tensors ([R, Kp] each)."""
    if not 1 <= len(specs) <= 2 or len({s.operand_mode for s in specs}) != 1:
        raise ValueError("stage: one or two specs of the same operand type")
    dev = specs[0].src.device
    with torch.cuda.device(dev):
        rc = _C.stage_launch([s.desc for s in specs], [s.out.data_ptr() for s in specs], [s.R for s in specs], [s.K for s in specs],
                             [s.out.shape[1] for s in specs], specs[0].operand_mode, _stream(specs[0].src))
    if rc != 0:
        raise RuntimeError(f"stage: rejected (code {rc})")
    return tuple(s.out for s in specs) if len(specs) > 1 else specs[0].out


def stage_gather(src, R, sizes, strides, s_row, operand_mode, base=0, rows_fast=None) -> torch.Tensor:
    return stage(spec_gather(src, R, sizes, strides, s_row, operand_mode, base, rows_fast))


def stage_matrix(x: torch.Tensor, operand_mode: int) -> torch.Tensor:
    return stage(spec_matrix(x, operand_mode))


def stage_im2col(x, kh, kw, out_hw, stride, pad, dil, operand_mode, in_dil=1, taps_are_rows=False) -> torch.Tensor:
    return stage(spec_im2col(x, kh, kw, out_hw, stride, pad, dil, operand_mode, in_dil, taps_are_rows))


def colsum_quantize(x2d: torch.Tensor, exp: int, man: int, rounding=ROUND_REF) -> torch.Tensor:
    """Q(x.sum(0)) of a row-major [rows, cols] fp32 view with unit column stride: the bias gradient in one launch."""
    _f32(x2d, "x"); _check_fmt(exp, man)
    if not x2d.is_cuda or x2d.dim() != 2 or x2d.stride(1) != 1:
        raise ValueError("colsum_quantize: 2-D CUDA tensor with unit inner stride expected")
    out = torch.empty(x2d.shape[1], dtype=torch.float32, device=x2d.device)
    with torch.cuda.device(x2d.device):
        _C.colsum_quantize_cuda(x2d.data_ptr(), x2d.shape[0], x2d.shape[1], x2d.stride(0) if x2d.shape[0] > 1 else x2d.shape[1],
                                out.data_ptr(), int(exp), int(man), rounding_code(rounding), _stream(x2d))
    return out


def mx_quantize(x: torch.Tensor, e5m2: bool = False):
    """fp32 [R, K] (row stride arbitrary) -> (codes uint8 [R, Kp], block scales uint8 atoms): MXFP8 with one
    UE8M0 power-of-two scale per 32 consecutive K-elements, laid out for tcgen05 (see mx_quantize_sm100.cu)."""
    _f32(x, "x")
    if not x.is_cuda or x.dim() != 2 or x.stride(1) != 1:
        raise ValueError("mx_quantize: 2-D CUDA tensor with unit inner stride expected")
    R, K = x.shape
    Kp = (K + 127) // 128 * 128
    codes = torch.empty(R, Kp, dtype=torch.uint8, device=x.device)
    sf = torch.zeros(((R + 127) // 128) * (Kp // 128) * 512, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        _C.mx_quantize_cuda(x.data_ptr(), x.stride(0), codes.data_ptr(), sf.data_ptr(), R, K, Kp, bool(e5m2), _stream(x))
    return codes, sf


def mx_dequantize(codes: torch.Tensor, sf: torch.Tensor, K: int, e5m2: bool = False) -> torch.Tensor:
    """Host-side model of what the tensor core sees: codes * 2^(scale-127), for tests."""
    R, Kp = codes.shape
    vals = codes.view(torch.float8_e5m2 if e5m2 else torch.float8_e4m3fn).float()
    r = torch.arange(R, device=codes.device)
    kb = torch.arange(Kp // 32, device=codes.device)
    atom = (r[:, None] // 128) * (Kp // 128) + (kb[None, :] // 4)
    off = atom * 512 + (r[:, None] % 32) * 16 + ((r[:, None] // 32) % 4) * 4 + (kb[None, :] % 4)
    scale = torch.exp2(sf[off].float() - 127.0)
    return (vals.view(R, Kp // 32, 32) * scale[:, :, None]).view(R, Kp)[:, :K]


def gemm_tc_mx(a_codes, a_sf, bt_codes, bt_sf, exp: int, man: int, kahan: bool = False, k_slice: int = 128, e5m2: bool = False,
               bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Block-scaled fp8 GEMM: C[M,N] = (A codes x A scales) @ (Bt codes x Bt scales)^T on tcgen05
    kind::mxf8f6f4.block_scale, accumulator re-quantised to (exp, man) after every K-slice."""
    _check_fmt(exp, man)
    M, Kp = a_codes.shape
    N = bt_codes.shape[0]
    if bt_codes.shape[1] != Kp or Kp % 128:
        raise ValueError("gemm_tc_mx: operands must share a K padded to a multiple of 128")
    out = torch.empty(M, N, dtype=torch.float32, device=a_codes.device)
    with torch.cuda.device(a_codes.device):
        rc = _C.gemm_tc_mx_cuda(a_codes.data_ptr(), a_sf.data_ptr(), bt_codes.data_ptr(), bt_sf.data_ptr(), out.data_ptr(), M, N, Kp,
                                int(exp), int(man), bool(kahan), int(k_slice), bool(e5m2), _ptr(bias), _stream(a_codes))
    if rc != 0:
        raise RuntimeError(f"gemm_tc_mx: kernel rejected the problem (code {rc})")
    return out


# --------------------------------------------------------- multi-tensor arenas --
def aps_fold(srcs: Sequence[torch.Tensor], out: torch.Tensor, seg_off: torch.Tensor, exp: int, man: int,
             use_aps: bool, kahan: bool, world_mult: float, chunks: Optional[torch.Tensor] = None,
             workspace: Optional[torch.Tensor] = None, shifts_out: Optional[torch.Tensor] = None,
             max_ctas: int = 0, rounding=ROUND_REF) -> torch.Tensor:
    """APS + rank-ordered (Kahan) quantised fold over len(srcs) flat arenas -> out."""
    _check_fmt(exp, man)
    r = rounding_code(rounding)
    for s in srcs:
        _f32(s, "src")
        if not s.is_contiguous() or s.numel() != out.numel() or s.device != out.device:
            raise ValueError("aps_fold: arenas must be flat, equal-sized and on one device")
    _f32(out, "out")
    T = seg_off.numel() - 1
    if shifts_out is not None and (shifts_out.dtype != torch.int32 or shifts_out.numel() < T):
        raise ValueError("shifts_out: int32[T] expected")
    if out.is_cuda:
        if chunks is None or workspace is None:
            raise ValueError("aps_fold: CUDA path needs chunks and workspace")
        if workspace.dtype != torch.int32 or workspace.numel() < 4 + 2 * T:
            raise ValueError("aps_fold: workspace must be int32[4 + 2*T_cap] with T_cap >= T")
        t_cap = (workspace.numel() - 4) // 2
        n_chunks = chunks.numel() // CHUNK_DTYPE.itemsize
        with torch.cuda.device(out.device):
            _C.aps_fold_cuda([s.data_ptr() for s in srcs], out.data_ptr(), chunks.data_ptr(), n_chunks, T, t_cap, int(exp),
                             int(man), r, bool(use_aps), bool(kahan), float(world_mult), workspace.data_ptr(),
                             _ptr(shifts_out), int(max_ctas), _stream(out))
    else:
        if seg_off.dtype != torch.int64 or seg_off.is_cuda:
            raise ValueError("seg_off: int64 CPU tensor expected")
        _C.aps_fold_cpu([s.data_ptr() for s in srcs], out.data_ptr(), seg_off.data_ptr(), T, int(exp), int(man), r,
                        bool(use_aps), bool(kahan), float(world_mult), _ptr(shifts_out))
    return out


def lars_step(master: torch.Tensor, model: Optional[torch.Tensor], grad: torch.Tensor, mom: torch.Tensor,
              seg_off: torch.Tensor, lr: float, momentum: float, wd: float, coeff: float = 0.001, eps: float = 0.0,
              use_lars: bool = True, nesterov: bool = False, wd_mult: Optional[torch.Tensor] = None,
              chunks: Optional[torch.Tensor] = None, tchunk_off: Optional[torch.Tensor] = None,
              lr_dev: Optional[torch.Tensor] = None, partials: Optional[torch.Tensor] = None,
              barrier_ws: Optional[torch.Tensor] = None, max_ctas: int = 0) -> None:
    for t, n in ((master, "master"), (grad, "grad"), (mom, "mom")):
        _f32(t, n)
        if not t.is_contiguous() or t.numel() != master.numel() or t.device != master.device:
            raise ValueError("lars_step: arenas must be flat, equal-sized and on one device")
    T = seg_off.numel() - 1
    if master.is_cuda:
        if chunks is None or tchunk_off is None or partials is None or barrier_ws is None:
            raise ValueError("lars_step: CUDA path needs chunks, tchunk_off, partials, barrier_ws")
        n_chunks = chunks.numel() // CHUNK_DTYPE.itemsize
        if partials.numel() < 2 * n_chunks:
            raise ValueError("lars_step: partials too small")
        with torch.cuda.device(master.device):
            _C.lars_step_cuda(master.data_ptr(), _ptr(model), grad.data_ptr(), mom.data_ptr(), chunks.data_ptr(),
                              tchunk_off.data_ptr(), n_chunks, T, _ptr(lr_dev), float(lr), float(momentum), float(wd),
                              float(coeff), float(eps), bool(use_lars), bool(nesterov), _ptr(wd_mult), partials.data_ptr(),
                              barrier_ws.data_ptr(), int(max_ctas), _stream(master))
    else:
        lrv = float(lr_dev.item()) if lr_dev is not None else float(lr)
        _C.lars_step_cpu(master.data_ptr(), _ptr(model), grad.data_ptr(), mom.data_ptr(), seg_off.data_ptr(), T, lrv,
                         float(momentum), float(wd), float(coeff), float(eps), bool(use_lars), bool(nesterov), _ptr(wd_mult))


def wire_roundtrip_cpu(x: torch.Tensor, exp: int, man: int) -> torch.Tensor:
    _f32(x, "x")
    x = x.contiguous()
    out = torch.empty_like(x)
    _C.wire_roundtrip_cpu(x.data_ptr(), out.data_ptr(), x.numel(), int(exp), int(man))
    return out


def topology() -> dict:
    import json
    return json.loads(_C.discover_topology_json())


def sm_count(device: int = -1) -> int:
    return _C.sm_count(device)
