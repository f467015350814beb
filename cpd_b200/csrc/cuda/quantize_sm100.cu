// quantize_sm100.cu — float_quantize for B200.
//
// Replaces the reference's float_kernel_nearest (float_kernel.cu:94-101: one
// element per thread, scalar 4-byte accesses, a data-dependent loop of up to 127
// multiplies per element, legacy default stream, int-sized) with a branch-light
// integer cast on the fp32 bit pattern, 128-bit streaming accesses, 4 independent
// vectors in flight per thread, a grid sized to the SM count, 64-bit indexing and
// the caller's stream.  Optional fused power-of-two scale (APS) and fused amax.
// Memory-bound: 8 B/element; roofline = measured HBM copy bandwidth.
#include "../api.h"
#include "../common/fpformat.h"
#include "../common/device_utils.cuh"
#include "format_dispatch.cuh"

namespace cpd {

constexpr int kQThreads = 256;
constexpr int kQUnroll = 4;   // float4 vectors per thread per iteration

// NaN payloads must pass through untouched (an FMUL would canonicalise them), so the
// scale is only applied when a non-zero APS shift was requested.
#define SC(v) (do_scale ? (v) * scale : (v))
template <class F>
__global__ void __launch_bounds__(kQThreads)
quantize_kernel(const float* in, float* out, long long n, F fmt, float scale, bool do_scale, bool vec_ok) {
  const long long tid = (long long)blockIdx.x * kQThreads + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * kQThreads;
  if (vec_ok) {
    const long long nvec = n >> 2;
    const float4* in4 = reinterpret_cast<const float4*>(in);
    float4* out4 = reinterpret_cast<float4*>(out);
    // main loop: kQUnroll coalesced vectors per thread, all loads issued before any use
    long long v = tid;
    for (; v + (kQUnroll - 1) * nthreads < nvec; v += kQUnroll * nthreads) {
      float4 x[kQUnroll];
#pragma unroll
      for (int j = 0; j < kQUnroll; ++j) x[j] = ld_stream_f4(in4 + v + j * nthreads);
#pragma unroll
      for (int j = 0; j < kQUnroll; ++j) {
        x[j].x = fmt.q(SC(x[j].x)); x[j].y = fmt.q(SC(x[j].y));
        x[j].z = fmt.q(SC(x[j].z)); x[j].w = fmt.q(SC(x[j].w));
        st_stream_f4(out4 + v + j * nthreads, x[j]);
      }
    }
    for (; v < nvec; v += nthreads) {
      float4 x = ld_stream_f4(in4 + v);
      x.x = fmt.q(SC(x.x)); x.y = fmt.q(SC(x.y));
      x.z = fmt.q(SC(x.z)); x.w = fmt.q(SC(x.w));
      st_stream_f4(out4 + v, x);
    }
    for (long long i = (nvec << 2) + tid; i < n; i += nthreads) out[i] = fmt.q(SC(in[i]));
  } else {
    for (long long i = tid; i < n; i += nthreads) out[i] = fmt.q(SC(in[i]));
  }
}

#undef SC

__global__ void __launch_bounds__(kQThreads)
amax_kernel(const float* __restrict__ in, long long n, unsigned* __restrict__ amax_bits, bool vec_ok) {
  __shared__ unsigned red[32];
  const long long tid = (long long)blockIdx.x * kQThreads + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * kQThreads;
  unsigned m = 0u;
  if (vec_ok) {
    const long long nvec = n >> 2;
    const uint4* in4 = reinterpret_cast<const uint4*>(in);
    for (long long v = tid; v < nvec; v += nthreads) {
      const uint4 x = ld_stream_u4(in4 + v);
      m = max(max(m, x.x & 0x7FFFFFFFu), max(x.y & 0x7FFFFFFFu, max(x.z & 0x7FFFFFFFu, x.w & 0x7FFFFFFFu)));
    }
    for (long long i = (nvec << 2) + tid; i < n; i += nthreads) m = max(m, __float_as_uint(in[i]) & 0x7FFFFFFFu);
  } else {
    for (long long i = tid; i < n; i += nthreads) m = max(m, __float_as_uint(in[i]) & 0x7FFFFFFFu);
  }
  m = block_max_u32(m, red);
  if (threadIdx.x == 0 && m) atomicMax(amax_bits, m);
}

static int grid_for(long long n_items, int sm_count, int waves) {
  long long blocks = (n_items + kQThreads - 1) / kQThreads;
  long long cap = (long long)sm_count * waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

int device_sm_count(int dev) {
  static int cached[64] = {0};
  if (dev < 0) cudaGetDevice(&dev);
  if (dev < 64 && cached[dev]) return cached[dev];
  int n = 148;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (dev < 64) cached[dev] = n;
  return n;
}

void quantize_cuda(const float* in, float* out, int64_t n, int e, int m, int rounding, int shift,
                   cudaStream_t st) {
  if (n <= 0) return;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
  const float scale = pow2i(shift);
  // 8 resident CTAs of 256 threads per SM (2048 threads) -> one full wave, grid-stride inside
  const int grid = grid_for(vec_ok ? (n + 3) / 4 : n, device_sm_count(-1), 8);
  CPD_DISPATCH_FORMAT(e, m, rounding, FMT, {
    quantize_kernel<decltype(FMT)><<<grid, kQThreads, 0, st>>>(in, out, (long long)n, FMT, scale, shift != 0, vec_ok);
  });
}

void amax_cuda(const float* in, int64_t n, unsigned* amax_bits, cudaStream_t st) {
  if (n <= 0) return;
  const bool vec_ok = (reinterpret_cast<uintptr_t>(in) & 15u) == 0;
  const int grid = grid_for(vec_ok ? (n + 3) / 4 : n, device_sm_count(-1), 8);
  amax_kernel<<<grid, kQThreads, 0, st>>>(in, (long long)n, amax_bits, vec_ok);
}

}  // namespace cpd
