// aps_fold_sm100.cu — single-GPU fused "APS + quantised ordered fold" over W local
// gradient arenas, one launch for a whole bucket of tensors.
//
// Covers (a) emulate_node: W = E micro-batch gradient slots (reference
// mix.py:251-285, main.py:178-202: per-parameter python loops with host syncs,
// clone lists and ~10 launches per parameter per slot) and (b) the W == 1 case of
// sum_gradients (dist_util.py:22-51 at world_size 1).  One persistent grid:
//   phase A  per-tensor max|g| over all sources (uint-bit atomicMax per tensor)
//   ---- grid barrier ----
//   phase B  shift = bias - ceil(log2(amax * mult)); res = fold_r Q(g_r * 2^shift)
//            (plain or Kahan, c in registers); out = res * 2^-shift
// Chunks are 16-element aligned pieces of one tensor (arena padding is zero), so
// every access is a 128-bit vector and one shift applies per chunk.
#include "../api.h"
#include "../common/fpformat.h"
#include "../common/device_utils.cuh"
#include "../common/wire_hw.cuh"
#include "format_dispatch.cuh"

namespace cpd {

constexpr int kFoldThreads = 256;
constexpr int kMaxLocalSrc = kMaxLocalSources;   // api.h: 64 (the reference's README runs --emulate-node 32)

struct SrcPtrs { const float* p[kMaxLocalSrc]; };

// Each thread owns up to kVecPerThread float4 vectors of a chunk, strided by the block size so a
// warp touches 512 contiguous bytes per vector; all of a thread's loads for one source are issued
// before the first use (4 x 16 B in flight per thread per source) — the kernel is latency-bound
// otherwise (Little: 148 SMs x resident warps x bytes in flight / ~1 us DRAM latency).
constexpr int kVecPerThread = 4;     // 4 x 256 threads x 4 floats = 4096 = CHUNK_ELEMS

template <class F, bool KAHAN>
__global__ void __launch_bounds__(kFoldThreads, 4)
aps_fold_local_kernel(SrcPtrs srcs, int W, float* out, const Chunk* __restrict__ chunks,
                      int n_chunks, int T, int T_cap, F fmt, bool use_aps, float world_mult,
                      unsigned* __restrict__ ws, int* __restrict__ shifts_out) {
  // ws layout: [0]=barrier count, [1]=barrier generation, [2]=launch counter, [4 .. 4+2*T_cap) = amax[2][T_cap]
  __shared__ unsigned red[32];
  __shared__ const float* s_src[kMaxLocalSrc];
  if (threadIdx.x < kMaxLocalSrc) s_src[threadIdx.x] = srcs.p[threadIdx.x];
  __syncthreads();
  unsigned* bar = ws;
  const unsigned launch = ws[2];
  unsigned* amax = ws + 4 + (launch & 1u) * T_cap;
  unsigned* amax_next = ws + 4 + ((launch & 1u) ^ 1u) * T_cap;

  // contiguous chunk range per CTA: consecutive chunks mostly belong to the same tensor, so the per-thread maxima are
  // carried across chunks and the block reduction + atomic happen only when the tensor changes (~1 per CTA for the
  // large tensors instead of 1 per chunk: two __syncthreads less per 16 KB) - and phase B revisits the SAME chunks,
  // which the CTA's SM has just pulled through L2
  const int per_cta = (n_chunks + gridDim.x - 1) / gridDim.x;
  const int c_lo = blockIdx.x * per_cta, c_hi = min(n_chunks, c_lo + per_cta);
  if (use_aps) {
    unsigned m = 0u;
    int cur_tensor = -1;
    for (int c = c_lo; c < c_hi; ++c) {
      const Chunk ck = chunks[c];
      const int nvec = ck.len >> 2;
      if (ck.tensor != cur_tensor) {
        if (cur_tensor >= 0) {
          m = block_max_u32(m, red);
          if (threadIdx.x == 0 && m) atomicMax(&amax[cur_tensor], m);
        }
        cur_tensor = ck.tensor;
        m = 0u;
      }
      for (int r = 0; r < W; ++r) {
        const uint4* base = reinterpret_cast<const uint4*>(s_src[r] + ck.start);
        uint4 x[kVecPerThread];
#pragma unroll
        for (int k = 0; k < kVecPerThread; ++k) {
          const int v = threadIdx.x + k * kFoldThreads;
          x[k] = v < nvec ? base[v] : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int k = 0; k < kVecPerThread; ++k)
          m = max(max(m, x[k].x & 0x7FFFFFFFu), max(x[k].y & 0x7FFFFFFFu, max(x[k].z & 0x7FFFFFFFu, x[k].w & 0x7FFFFFFFu)));
      }
    }
    if (cur_tensor >= 0) {
      m = block_max_u32(m, red);
      if (threadIdx.x == 0 && m) atomicMax(&amax[cur_tensor], m);
    }
    grid_barrier(bar);
  }

  for (int c = c_lo; c < c_hi; ++c) {
    const Chunk ck = chunks[c];
    const int nvec = ck.len >> 2;
    int shift = 0;
    if (use_aps) {
      const unsigned ab = amax[ck.tensor];
      const float me = ab >= 0x7F800000u ? INFINITY : aps_local_exponent(__uint_as_float(ab), world_mult);
      shift = aps_shift_from_exponent(me, fmt.exp_bits());
      if (shifts_out && threadIdx.x == 0) shifts_out[ck.tensor] = shift;
    }
    const float sc = pow2i(shift), inv = pow2i(-shift);
    float s[kVecPerThread][4], cc[KAHAN ? kVecPerThread : 1][4];
#pragma unroll
    for (int k = 0; k < kVecPerThread; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) { s[k][j] = 0.f; if (KAHAN) cc[KAHAN ? k : 0][j] = 0.f; }
    for (int r = 0; r < W; ++r) {
      const float4* base = reinterpret_cast<const float4*>(s_src[r] + ck.start);
      float4 g4[kVecPerThread];
#pragma unroll
      for (int k = 0; k < kVecPerThread; ++k) {
        const int v = threadIdx.x + k * kFoldThreads;
        g4[k] = v < nvec ? base[v] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int k = 0; k < kVecPerThread; ++k) {
        float g[4] = {g4[k].x, g4[k].y, g4[k].z, g4[k].w};
        if (use_aps) {
          // the scaled casts make every value a member of the format, so every later cast acts on an exact sum of
          // members: hardware-assisted exact path for the fp8-shaped formats, 4 casts per guard (wire_hw.cuh); the
          // scaled cast itself goes through the conversion hardware too (pretruncate + F2FP)
          FoldOps<F>::template qscaled_n<4>(fmt, g, sc);
          if (W == 1) {
            // one source: Q(0 + v) == 0 + v and Kahan's (y, t, c) == (v, 0 + v, 0) for a member v - nothing left to cast
#pragma unroll
            for (int j = 0; j < 4; ++j) s[k][j] = 0.f + g[j];          // "+ 0": the fold turns -0 into +0
          } else if (KAHAN) fold_kahan_members_n(fmt, s[k], cc[KAHAN ? k : 0], g);
          else fold_plain_members_n(fmt, s[k], g);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (KAHAN) fold_kahan(fmt, s[k][j], cc[KAHAN ? k : 0][j], g[j]);
            else fold_plain(fmt, s[k][j], g[j]);
          }
        }
      }
    }
    float4* obase = reinterpret_cast<float4*>(out + ck.start);
#pragma unroll
    for (int k = 0; k < kVecPerThread; ++k) {
      const int v = threadIdx.x + k * kFoldThreads;
      if (v < nvec) {
        float4 o;
        if (use_aps) { o.x = s[k][0] * inv; o.y = s[k][1] * inv; o.z = s[k][2] * inv; o.w = s[k][3] * inv; }
        else { o.x = s[k][0]; o.y = s[k][1]; o.z = s[k][2]; o.w = s[k][3]; }
        obase[v] = o;
      }
    }
  }

  // prepare the other amax buffer for the next launch and bump the launch counter.
  // Safe without another barrier: nobody touches amax_next in this launch, and ws[2]
  // was read by every CTA before the grid barrier.
  if (use_aps && blockIdx.x == 0) {
    for (int t = threadIdx.x; t < T_cap; t += kFoldThreads) amax_next[t] = 0u;
    __syncthreads();
    if (threadIdx.x == 0) ws[2] = launch + 1u;
  }
}

void aps_fold_local_cuda(const float* const* srcs_host, int W, float* out, const Chunk* chunks,
                         int n_chunks, int T, int T_cap, int e, int m, int rounding, bool use_aps, bool kahan, float world_mult,
                         unsigned* workspace, int* shifts_out, int max_ctas, cudaStream_t st) {
  if (n_chunks <= 0 || W <= 0) return;
  SrcPtrs sp;
  for (int r = 0; r < kMaxLocalSrc; ++r) sp.p[r] = r < W ? srcs_host[r] : nullptr;
  int grid = max_ctas > 0 ? max_ctas : device_sm_count(-1) * 4;   // 4 co-resident CTAs / SM (launch bounds)
  if (grid > n_chunks) grid = n_chunks;
  CPD_DISPATCH_FORMAT(e, m, rounding, FMT, {
    if (kahan)
      aps_fold_local_kernel<decltype(FMT), true><<<grid, kFoldThreads, 0, st>>>(
          sp, W, out, chunks, n_chunks, T, T_cap, FMT, use_aps, world_mult, workspace, shifts_out);
    else
      aps_fold_local_kernel<decltype(FMT), false><<<grid, kFoldThreads, 0, st>>>(
          sp, W, out, chunks, n_chunks, T, T_cap, FMT, use_aps, world_mult, workspace, shifts_out);
  });
}

}  // namespace cpd
