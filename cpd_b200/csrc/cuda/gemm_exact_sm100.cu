// gemm_exact_sm100.cu — SIMT GEMM with the reference's exact per-MAC semantics:
// strictly sequential k, every product and every Kahan intermediate cast to
// (exp,man) (reference tvm_gemm, float_kernel.cu:103-340).
// This is the parity / small-shape path; the throughput path is the tcgen05
// kernel (gemm_tcgen05_sm100a.cu) with per-K-slice re-quantisation.
//
// Differences from the reference kernel by design: 64x64 tile / 256 threads / 4x4
// micro-tile (16x the reuse of its 16x16 / 64-thread / 2x2 tile), arbitrary 2-D
// strides (no materialised transposes for A^T or B^T), any M,N,K with correct
// remainder handling (the reference's M-remainder branch uses an uninitialised
// compensation array and no barriers), compile-time specialised casts, caller's stream.
#include "../api.h"
#include "../common/fpformat.h"
#include "format_dispatch.cuh"

namespace cpd {

constexpr int GE_BM = 64, GE_BN = 64, GE_BK = 16, GE_THREADS = 256;

template <class F, bool KAHAN>
__global__ void __launch_bounds__(GE_THREADS)
gemm_exact_kernel(const float* __restrict__ A, long long sa0, long long sa1, const float* __restrict__ B,
                  long long sb0, long long sb1, float* __restrict__ C, long long ldc, int M, int N, int K,
                  F fmt) {
  __shared__ float As[GE_BK][GE_BM + 4];
  __shared__ float Bs[GE_BK][GE_BN + 4];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;   // 16 x 16 threads, 4x4 outputs each
  const int m0 = blockIdx.y * GE_BM, n0 = blockIdx.x * GE_BN;
  float s[4][4], c[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { s[i][j] = 0.f; c[i][j] = 0.f; }

  for (int k0 = 0; k0 < K; k0 += GE_BK) {
    // cooperative tile loads (zero-filled out of range)
    for (int idx = threadIdx.x; idx < GE_BK * GE_BM; idx += GE_THREADS) {
      int kk, mm;
      if (sa1 == 1) { kk = idx % GE_BK; mm = idx / GE_BK; } else { mm = idx % GE_BM; kk = idx / GE_BM; }
      const int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < K) ? A[gm * sa0 + gk * sa1] : 0.f;
    }
    for (int idx = threadIdx.x; idx < GE_BK * GE_BN; idx += GE_THREADS) {
      int kk, nn;
      if (sb1 == 1) { nn = idx % GE_BN; kk = idx / GE_BN; } else { kk = idx % GE_BK; nn = idx / GE_BK; }
      const int gn = n0 + nn, gk = k0 + kk;
      Bs[kk][nn] = (gn < N && gk < K) ? B[gk * sb0 + gn * sb1] : 0.f;
    }
    __syncthreads();
    const int kmax = min(GE_BK, K - k0);
    for (int kk = 0; kk < kmax; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float p = fmt.q(__fmul_rn(a[i], b[j]));   // product rounded to fp32 first, never fused
          if (KAHAN) fold_kahan(fmt, s[i][j], c[i][j], p);
          else fold_plain(fmt, s[i][j], p);
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn < N) C[gm * ldc + gn] = s[i][j];
    }
  }
}

void gemm_exact_cuda(const float* A, int64_t sa0, int64_t sa1, const float* B, int64_t sb0, int64_t sb1,
                     float* C, int64_t ldc, int M, int N, int K, int e, int m, int rounding, bool kahan,
                     cudaStream_t st) {
  if (M <= 0 || N <= 0) return;
  dim3 grid((N + GE_BN - 1) / GE_BN, (M + GE_BM - 1) / GE_BM);
  CPD_DISPATCH_FORMAT(e, m, rounding, FMT, {
    if (kahan)
      gemm_exact_kernel<decltype(FMT), true><<<grid, GE_THREADS, 0, st>>>(A, sa0, sa1, B, sb0, sb1, C, ldc, M, N, K, FMT);
    else
      gemm_exact_kernel<decltype(FMT), false><<<grid, GE_THREADS, 0, st>>>(A, sa0, sa1, B, sb0, sb1, C, ldc, M, N, K, FMT);
  });
}

}  // namespace cpd
