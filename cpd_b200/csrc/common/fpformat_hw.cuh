// fpformat_hw.cuh — IEEE-rounding casts that map onto Blackwell's conversion hardware.
//
// For (5,10) == fp16, (8,7) == bf16, (4,3) ~ E4M3 and (5,2) ~ E5M2 the cast
// fp32 -> (exp,man) -> fp32 with IEEE RNE, target subnormals and overflow-to-Inf is a pair of
// F2F/F2FP instructions instead of ~10 integer ops.  CPD's (4,3)/(5,2) reserve the all-ones
// exponent (max normal 240 / 57344, then Inf) whereas OCP E4M3 extends to 448 and E5M2's
// satfinite convert clamps, so the top of the range is patched with one compare + select.
// These policies implement `rounding="ieee"` exactly (same results as quantize_ieee in
// fpformat.h); they are NOT bit-compatible with the reference's quirks A/B and are therefore
// only used where the semantics are defined as IEEE: the tensor-core GEMM's per-K-slice requant.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include "fpformat.h"

namespace cpd {

struct FmtHwF16 {      // (5, 10)
  __device__ __forceinline__ float q(float x) const { return __half2float(__float2half_rn(x)); }
  __device__ __forceinline__ int exp_bits() const { return 5; }
};

struct FmtHwBf16 {     // (8, 7)
  __device__ __forceinline__ float q(float x) const { return __bfloat162float(__float2bfloat16_rn(x)); }
  __device__ __forceinline__ int exp_bits() const { return 8; }
};

struct FmtHwE4M3 {     // (4, 3): values below 248 coincide with OCP E4M3; >= 248 rounds to Inf
  __device__ __forceinline__ float q(float x) const {
    const __nv_fp8_storage_t c = __nv_cvt_float_to_fp8(x, __NV_SATFINITE, __NV_E4M3);
    const float r = __half2float(__half(__nv_cvt_fp8_to_halfraw(c, __NV_E4M3)));
    return fabsf(x) >= 248.0f ? copysignf(__int_as_float(0x7F800000), x) : r;     // NaN compares false -> r (NaN)
  }
  __device__ __forceinline__ int exp_bits() const { return 4; }
};

struct FmtHwE5M2 {     // (5, 2): >= 61440 rounds to Inf
  __device__ __forceinline__ float q(float x) const {
    const __nv_fp8_storage_t c = __nv_cvt_float_to_fp8(x, __NV_SATFINITE, __NV_E5M2);
    const float r = __half2float(__half(__nv_cvt_fp8_to_halfraw(c, __NV_E5M2)));
    return fabsf(x) >= 61440.0f ? copysignf(__int_as_float(0x7F800000), x) : r;
  }
  __device__ __forceinline__ int exp_bits() const { return 5; }
};

struct FmtHwF32 {      // (8, 23): identity
  __device__ __forceinline__ float q(float x) const { return x; }
  __device__ __forceinline__ int exp_bits() const { return 8; }
};

}  // namespace cpd

// IEEE-rounding dispatch: hardware policies where they exist, the runtime software path otherwise.
#define CPD_DISPATCH_FORMAT_IEEE_HW(E, M, NAME, ...)                                      \
  do {                                                                                    \
    const int _e = (E), _m = (M);                                                         \
    if (_e == 5 && _m == 10) { ::cpd::FmtHwF16 NAME; __VA_ARGS__ }                        \
    else if (_e == 8 && _m == 7) { ::cpd::FmtHwBf16 NAME; __VA_ARGS__ }                   \
    else if (_e == 4 && _m == 3) { ::cpd::FmtHwE4M3 NAME; __VA_ARGS__ }                   \
    else if (_e == 5 && _m == 2) { ::cpd::FmtHwE5M2 NAME; __VA_ARGS__ }                   \
    else if (_e == 8 && _m == 23) { ::cpd::FmtHwF32 NAME; __VA_ARGS__ }                   \
    else { ::cpd::FmtR NAME{::cpd::make_format(_e, _m, ::cpd::ROUND_IEEE)}; __VA_ARGS__ } \
  } while (0)
