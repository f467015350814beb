// mx_quantize_sm100.cu — operand staging for the block-scaled fp8 GEMM (kind::mxf8f6f4.block_scale).
//
// fp32 [R, K] row-major  ->  fp8 codes [R, Kp] (E4M3 or E5M2, Kp = K rounded up to 128, zero padded)
//                          + UE8M0 block scales, one per 32 consecutive K-elements of a row, written
//                            directly in the 512-byte "scale-factor atom" layout that tcgen05.cp moves
//                            into TMEM: atom (R/128, K/128) holds rows r (128) x 4 k-blocks at byte
//                            (r % 32) * 16 + (r / 32 % 4) * 4 + kblock.
// The scale of a block is the smallest power of two s with max|x| / s <= max finite code
// (448 for E4M3, 57344 for E5M2), so nothing saturates; x / s is converted with the hardware
// F2FP.SATFINITE convert (RNE).  This is the APS idea (shift by a power of two taken from the
// maximum exponent) applied per 32-element block instead of per tensor.
// One warp converts 128 consecutive elements of a row: 4 elements per lane, 8 lanes per block.
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "../api.h"

namespace cpd {

template <bool E5M2>
__global__ void __launch_bounds__(256)
mx_quantize_kernel(const float* __restrict__ in, uint8_t* __restrict__ codes, uint8_t* __restrict__ sf, int R, int K, int Kp,
                   long long ld_in) {
  const int warps_per_row = Kp >> 7;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= (long long)R * warps_per_row) return;
  const int r = (int)(gw / warps_per_row), kb = (int)(gw % warps_per_row);
  const int k0 = kb * 128 + lane * 4;
  float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* row = in + (long long)r * ld_in;
  if (k0 + 3 < K && ((reinterpret_cast<uintptr_t>(row + k0) & 15u) == 0)) {
    x = *reinterpret_cast<const float4*>(row + k0);
  } else {
    if (k0 + 0 < K) x.x = row[k0 + 0];
    if (k0 + 1 < K) x.y = row[k0 + 1];
    if (k0 + 2 < K) x.z = row[k0 + 2];
    if (k0 + 3 < K) x.w = row[k0 + 3];
  }
  float amax = fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w)));
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));     // over the 8 lanes of a block
  // smallest power of two s.t. amax / 2^e <= max code: e = ceil(log2(amax / maxcode)) from the bit pattern
  constexpr float kMaxCode = E5M2 ? 57344.0f : 448.0f;
  int e = -127;
  if (amax > 0.f && amax < INFINITY) {
    const float q = amax / kMaxCode;
    const uint32_t qb = __float_as_uint(q);
    e = (int)((qb >> 23) & 0xFF) - 127 + ((qb & 0x7FFFFFu) ? 1 : 0);
    if ((qb >> 23) == 0) e = -126;
  }
  e = max(-127, min(127, e));
  const float inv = __uint_as_float(e == 127 ? 0x00400000u : ((uint32_t)(127 - e) << 23));   // 2^-e (denormal bit pattern for e = 127)
  const __nv_fp8_interpretation_t interp = E5M2 ? __NV_E5M2 : __NV_E4M3;
  const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(x.x * inv, x.y * inv), __NV_SATFINITE, interp);
  const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(x.z * inv, x.w * inv), __NV_SATFINITE, interp);
  *reinterpret_cast<uint32_t*>(codes + (long long)r * Kp + k0) = lo | (hi << 16);
  // the 4 block scales of this warp sit in lanes 0, 8, 16, 24 -> one 32-bit store by lane 0
  const uint32_t my = (uint32_t)(e + 127);
  const uint32_t s0 = __shfl_sync(0xffffffffu, my, 0), s1 = __shfl_sync(0xffffffffu, my, 8);
  const uint32_t s2 = __shfl_sync(0xffffffffu, my, 16), s3 = __shfl_sync(0xffffffffu, my, 24);
  if (lane == 0) {
    const long long atom = ((long long)(r >> 7) * warps_per_row + kb) * 512;
    const int within = (r & 31) * 16 + ((r >> 5) & 3) * 4;
    *reinterpret_cast<uint32_t*>(sf + atom + within) = s0 | (s1 << 8) | (s2 << 16) | (s3 << 24);
  }
}

// codes: [R, Kp] bytes; sf: [ceil(R/128) * (Kp/128) * 512] bytes, must be zero-initialised (rows beyond R keep scale 2^-127)
void mx_quantize_cuda(const float* in, long long ld_in, uint8_t* codes, uint8_t* sf, int R, int K, int Kp, bool e5m2, cudaStream_t st) {
  if (R <= 0 || K <= 0) return;
  const long long warps = (long long)R * (Kp >> 7);
  const int threads = 256;
  const long long blocks = (warps * 32 + threads - 1) / threads;
  if (e5m2) mx_quantize_kernel<true><<<(unsigned)blocks, threads, 0, st>>>(in, codes, sf, R, K, Kp, ld_in);
  else mx_quantize_kernel<false><<<(unsigned)blocks, threads, 0, st>>>(in, codes, sf, R, K, Kp, ld_in);
}

}  // namespace cpd
