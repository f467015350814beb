// operand_stage_sm100.cu — ONE kernel that turns an fp32 tensor into a tensor-core operand.
//
// The reference's custom-precision layers glue their GEMM together with torch ops
// (CPDtorch/quant/quant_module.py:26-52,115-139 + quant_function.py:95-97): F.unfold (an fp32 im2col 9x the size of
// the activations), transposes materialised by .contiguous(), a zero-filled output, a separate bias add.  Round 1 of
// this repo kept that shape (pad + .to(dtype) + .t().contiguous() around the tcgen05 kernel).  This kernel replaces
// all of it on the operand side: it gathers the logical matrix
//        out[r][k],  r < R rows,  k < K reduction index        (row-major, row stride Kp >= K, K-major = what TMA wants)
// from fp32 memory through one of two address generators, casts to the operand type of the tensor-core kind
// (bf16 / fp32-as-tf32 / e4m3 / e5m2, saturating: an fp8 cast of |x| > 448 must not become NaN) and writes 16-byte
// vectors.  Address generators:
//   GATHER   k = (k1, k2, k3) with sizes (n1, n2, n3) and arbitrary (also negative) element strides, rows with their own
//            stride: plain / transposed 2-D matrices, [O,C,KH,KW] weights viewed as [O,(kh,kw,c)], spatially flipped
//            weights for the data-gradient, dy^T for the weight-gradient — no materialised transposes.
//   IM2COL   rows <-> output pixels (b, oh, ow), k <-> filter taps (kh, kw, c) of a 2-D convolution over an activation
//            tensor with arbitrary (NCHW or NHWC) strides, stride / padding / dilation and an optional input dilation
//            (zero-upsampled dy of a strided convolution's data-gradient); `taps_are_rows` swaps the roles (the
//            weight-gradient's B operand is the TRANSPOSED im2col matrix).  Out-of-image taps are zeros.
// Tiles of 64 x 64 go through shared memory, so whichever of the two logical dimensions is contiguous in the source
// is read coalesced and the destination is always written in whole 16-byte vectors along k.
// K tails are NOT padded by anyone: the GEMM's TMA descriptor declares the true K and zero-fills out of bounds.
// A second small kernel does the bias gradient, column sums of dy cast to (exp, man): quant_module.py:49-50.
#include <cuda_bf16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "../api.h"
#include "../common/fpformat.h"
#include "format_dispatch.cuh"

namespace cpd {

namespace {

constexpr int kTile = 64;
constexpr int kStageThreads = 256;

// Index arithmetic is hoisted out of the element loop: the decomposition of a row index (pixel -> (b, oh, ow) or tap ->
// (kh, kw, c)) and of a column index is done ONCE per tile row / tile column (128 decompositions per 4096-element tile,
// in 32 bits) and parked in shared memory; an element then costs two small table reads, two adds and a bounds test.
// (The first version decomposed per element with 64-bit divisions: ~150 instructions per element, 0.14 of HBM bandwidth.)
struct StageTabs {
  long long ra[kTile], ca[kTile];      // address contributions of the tile's rows / columns (elements)
  int rh[kTile], rw[kTile];            // IM2COL: spatial contribution of the row (pixel origin or tap offset), or a poison value
  int ch[kTile], cw[kTile];            // IM2COL: spatial contribution of the column
};

constexpr int kOut = -(1 << 28);       // spatial coordinate of an out-of-range row / column: fails every bounds test

__device__ __forceinline__ void stage_decompose(const StageDesc& d, bool is_row, unsigned idx, bool in_range, long long& addr, int& h, int& w) {
  addr = 0; h = 0; w = 0;
  if (!in_range) { h = kOut; w = kOut; return; }
  if (d.mode == 0) {
    if (is_row) { addr = d.base + (long long)idx * d.s_row; return; }
    const unsigned n23 = (unsigned)d.n2 * (unsigned)d.n3;
    const unsigned k1 = idx / n23, rem = idx - k1 * n23;
    const unsigned k2 = rem / (unsigned)d.n3, k3 = rem - k2 * (unsigned)d.n3;
    addr = (long long)k1 * d.s1 + (long long)k2 * d.s2 + (long long)k3 * d.s3;
    return;
  }
  const bool is_pixel = is_row != (d.taps_are_rows != 0);
  if (is_pixel) {
    const unsigned ohw = (unsigned)(d.OH * d.OW);
    const unsigned b = idx / ohw, prem = idx - b * ohw;
    const int oh = (int)(prem / (unsigned)d.OW), ow = (int)(prem - (unsigned)oh * (unsigned)d.OW);
    addr = d.base + (long long)b * d.sb;
    h = oh * d.stride - d.pad; w = ow * d.stride - d.pad;
  } else {
    const unsigned kwc = (unsigned)(d.KW * d.C);
    const int kh = (int)(idx / kwc);
    const unsigned trem = idx - (unsigned)kh * kwc;
    const int kw = (int)(trem / (unsigned)d.C), c = (int)(trem - (unsigned)kw * (unsigned)d.C);
    addr = (long long)c * d.sc;
    h = kh * d.dil; w = kw * d.dil;
  }
}

__device__ __forceinline__ float stage_fetch(const StageDesc& d, const StageTabs& t, int rr, int kk) {
  if (d.mode == 0) {
    if (t.rh[rr] == kOut || t.ch[kk] == kOut) return 0.f;
    return d.src[t.ra[rr] + t.ca[kk]];
  }
  int ih = t.rh[rr] + t.ch[kk], iw = t.rw[rr] + t.cw[kk];          // an out-of-range row / column makes both hugely negative
  if (ih < 0 || iw < 0) return 0.f;
  if (d.in_dil > 1) {                                               // input dilation: only every in_dil-th coordinate carries data
    if (ih % d.in_dil || iw % d.in_dil) return 0.f;
    ih /= d.in_dil; iw /= d.in_dil;
  }
  if (ih >= d.H || iw >= d.W) return 0.f;
  return d.src[t.ra[rr] + t.ca[kk] + (long long)ih * d.sh + (long long)iw * d.sw];
}

// Four consecutive elements along the source-contiguous ("fast") dimension in ONE 16-byte load (desc.vec4, decided on the
// host: unit stride along that dimension, every other stride / offset / extent a multiple of 4, 16-byte aligned base).
// The four share their spatial coordinates (same filter tap, channels c..c+3; or same pixel), so one bounds test serves.
__device__ __forceinline__ float4 stage_fetch4(const StageDesc& d, const StageTabs& t, int rr, int kk) {
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  if (d.mode == 0) {
    if (t.rh[rr] == kOut || t.ch[kk] == kOut) return zero;
    return *reinterpret_cast<const float4*>(d.src + t.ra[rr] + t.ca[kk]);
  }
  int ih = t.rh[rr] + t.ch[kk], iw = t.rw[rr] + t.cw[kk];
  if (ih < 0 || iw < 0) return zero;
  if (d.in_dil > 1) {
    if (ih % d.in_dil || iw % d.in_dil) return zero;
    ih /= d.in_dil; iw /= d.in_dil;
  }
  if (ih >= d.H || iw >= d.W) return zero;
  return *reinterpret_cast<const float4*>(d.src + t.ra[rr] + t.ca[kk] + (long long)ih * d.sh + (long long)iw * d.sw);
}

// 8 consecutive k of one row -> operand dtype, one vector store
template <int OP>
__device__ __forceinline__ void stage_store8(void* dst_row, int k, const float (&v)[8]) {
  if (OP == 0) {                                                       // bf16
    uint4 o;
    __nv_bfloat162 h;
    h = __floats2bfloat162_rn(v[0], v[1]); o.x = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(v[2], v[3]); o.y = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(v[4], v[5]); o.z = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(v[6], v[7]); o.w = *reinterpret_cast<uint32_t*>(&h);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(dst_row) + k) = o;
  } else if (OP == 1) {                                                // fp32 (consumed as tf32)
    float* p = reinterpret_cast<float*>(dst_row) + k;
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  } else {                                                             // e4m3 / e5m2, saturating
    constexpr __nv_fp8_interpretation_t I = OP == 2 ? __NV_E4M3 : __NV_E5M2;
    uint2 o;
    o.x = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(v[0], v[1]), __NV_SATFINITE, I) |
          ((uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(v[2], v[3]), __NV_SATFINITE, I) << 16);
    o.y = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(v[4], v[5]), __NV_SATFINITE, I) |
          ((uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(v[6], v[7]), __NV_SATFINITE, I) << 16);
    *reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(dst_row) + k) = o;
  }
}

// One launch stages BOTH operands of a GEMM: the tile index space is the concatenation of operand 0's and operand 1's
// tiles (a GEMM then costs two launches in total: this one and the tensor-core kernel).
struct StagePair {
  StageDesc d[2];
  void* dst[2];
  long long R[2], K[2], Kp[2];
  int n;
};

template <int OP>
__global__ void __launch_bounds__(kStageThreads)
operand_stage_kernel(const StagePair sp) {
  __shared__ float tile[kTile][kTile + 1];
  __shared__ StageTabs tabs;
  constexpr int ES = OP == 0 ? 2 : (OP == 1 ? 4 : 1);
  const long long tk0 = (sp.Kp[0] + kTile - 1) / kTile;
  const long long nt0 = ((sp.R[0] + kTile - 1) / kTile) * tk0;
  const long long tk1 = sp.n > 1 ? (sp.Kp[1] + kTile - 1) / kTile : 1;
  const long long nt1 = sp.n > 1 ? ((sp.R[1] + kTile - 1) / kTile) * tk1 : 0;
  for (long long tt = blockIdx.x; tt < nt0 + nt1; tt += gridDim.x) {
    const int which = tt >= nt0 ? 1 : 0;
    const StageDesc& d = sp.d[which];
    void* dst = sp.dst[which];
    const long long R = sp.R[which], K = sp.K[which], Kp = sp.Kp[which];
    const long long t = which ? tt - nt0 : tt, tiles_k = which ? tk1 : tk0;
    const long long r0 = (t / tiles_k) * kTile, k0 = (t % tiles_k) * kTile;
    // index tables of this tile: threads 0..63 take the rows, 64..127 the columns
    if (threadIdx.x < 2 * kTile) {
      const bool is_row = threadIdx.x < kTile;
      const int j = is_row ? threadIdx.x : threadIdx.x - kTile;
      const long long idx = (is_row ? r0 : k0) + j;
      long long addr; int h, w;
      stage_decompose(d, is_row, (unsigned)idx, idx < (is_row ? R : K), addr, h, w);
      if (is_row) { tabs.ra[j] = addr; tabs.rh[j] = h; tabs.rw[j] = w; }
      else { tabs.ca[j] = addr; tabs.ch[j] = h; tabs.cw[j] = w; }
    }
    __syncthreads();
    // load phase: the source-contiguous dimension runs across the threads of a warp
    if (d.vec4) {
      // 16 lanes x 16 B cover the 64 fast indices of one slow index: 4 elements per load instead of 1 (the scalar loop
      // was issue-bound: ~30 instructions per element, ncu r2t4_prof_operand_stage_kernel)
#pragma unroll
      for (int i = 0; i < kTile * kTile / 4 / kStageThreads; ++i) {
        const int lin = i * kStageThreads + threadIdx.x;
        const int f0 = (lin % (kTile / 4)) * 4, slow = lin / (kTile / 4);
        if (d.rows_fast) {
          const float4 v = stage_fetch4(d, tabs, f0, slow);
          tile[f0][slow] = v.x; tile[f0 + 1][slow] = v.y; tile[f0 + 2][slow] = v.z; tile[f0 + 3][slow] = v.w;
        } else {
          const float4 v = stage_fetch4(d, tabs, slow, f0);
          tile[slow][f0] = v.x; tile[slow][f0 + 1] = v.y; tile[slow][f0 + 2] = v.z; tile[slow][f0 + 3] = v.w;
        }
      }
    } else {
#pragma unroll 4
      for (int i = 0; i < kTile * kTile / kStageThreads; ++i) {
        const int lin = i * kStageThreads + threadIdx.x;
        const int fast = lin % kTile, slow = lin / kTile;
        const int rr = d.rows_fast ? fast : slow, kk = d.rows_fast ? slow : fast;
        tile[rr][kk] = stage_fetch(d, tabs, rr, kk);
      }
    }
    __syncthreads();
    // store phase: 8 consecutive k per thread, rows across the rest
#pragma unroll
    for (int j = 0; j < kTile * kTile / 8 / kStageThreads; ++j) {
      const int lin = j * kStageThreads + threadIdx.x;
      const int rr = lin / (kTile / 8), kg = (lin % (kTile / 8)) * 8;
      const long long r = r0 + rr, k = k0 + kg;
      if (r < R && k < Kp) {
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = tile[rr][kg + q];
        stage_store8<OP>(reinterpret_cast<uint8_t*>(dst) + r * Kp * ES, (int)k, v);
      }
    }
    __syncthreads();
  }
}

// grad_bias[j] = Q(sum_i dy[i][j])  (quant_module.py:49-50: float_quantize(grad_output.sum(0)))
template <class F>
__global__ void __launch_bounds__(256)
colsum_quantize_kernel(const float* __restrict__ x, long long rows, int cols, long long ld, float* __restrict__ out, F fmt) {
  __shared__ float part[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ry = threadIdx.x >> 5;
  float s = 0.f;
  if (c < cols)
    for (long long r = ry; r < rows; r += 8) s += x[r * ld + c];
  part[ry][threadIdx.x & 31] = s;
  __syncthreads();
  if (ry == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x & 31];
    out[c] = fmt.q(t);
  }
}

}  // namespace

// Can groups of 4 along the source-contiguous dimension be fetched with one aligned 16-byte load?
static bool stage_can_vec4(const StageDesc& d, long long R, long long K) {
  auto m4 = [](long long v) { return (v & 3) == 0; };
  if (reinterpret_cast<uintptr_t>(d.src) & 15u) return false;
  if (!m4(d.base)) return false;
  if (d.mode == 0) {
    if (d.rows_fast) {                                   // rows contiguous: r, r+1, r+2, r+3 at one k
      if (d.s_row != 1 || !m4(R)) return false;
      return m4(d.s1) && (d.n2 == 1 || m4(d.s2)) && (d.n3 == 1 || m4(d.s3));
    }
    if (!m4(K) || !m4(d.s_row)) return false;
    if (d.n2 == 1 && d.n3 == 1) return d.s1 == 1;        // plain 2-D, K contiguous
    return d.s3 == 1 && m4(d.n3) && m4(d.s1) && m4(d.s2);    // 3-level gather with a contiguous innermost level
  }
  // IM2COL: channels contiguous (channels_last) and they are the fast dimension of this staging
  if (d.sc != 1 || !m4(d.C) || !m4(d.sb) || !m4(d.sh) || !m4(d.sw)) return false;
  return (d.rows_fast != 0) == (d.taps_are_rows != 0);
}

int operand_stage2_cuda(const StageDesc* descs, void* const* dsts, const long long* R, const long long* K, const long long* Kp,
                        int n, int operand_mode, cudaStream_t st) {
  if (n < 1 || n > 2) return -3;
  StagePair sp{};
  sp.n = n;
  long long n_tiles = 0;
  for (int i = 0; i < n; ++i) {
    if (R[i] <= 0 || K[i] <= 0 || R[i] >= (1LL << 31) || K[i] >= (1LL << 31)) return -4;      // 32-bit index arithmetic
    if (Kp[i] < K[i] || Kp[i] % 8) return -1;                       // rows are written in 8-element vectors
    sp.d[i] = descs[i]; sp.dst[i] = dsts[i]; sp.R[i] = R[i]; sp.K[i] = K[i]; sp.Kp[i] = Kp[i];
    sp.d[i].vec4 = stage_can_vec4(descs[i], R[i], K[i]) ? 1 : 0;
    n_tiles += ((R[i] + kTile - 1) / kTile) * ((Kp[i] + kTile - 1) / kTile);
  }
  const long long cap = (long long)device_sm_count(-1) * 8;
  const int grid = (int)(n_tiles < cap ? n_tiles : cap);
  switch (operand_mode) {
    case 0: operand_stage_kernel<0><<<grid, kStageThreads, 0, st>>>(sp); break;
    case 1: operand_stage_kernel<1><<<grid, kStageThreads, 0, st>>>(sp); break;
    case 2: operand_stage_kernel<2><<<grid, kStageThreads, 0, st>>>(sp); break;
    case 3: operand_stage_kernel<3><<<grid, kStageThreads, 0, st>>>(sp); break;
    default: return -2;
  }
  return 0;
}

int operand_stage_cuda(const StageDesc& d, void* dst, long long R, long long K, long long Kp, int operand_mode, cudaStream_t st) {
  if (R <= 0 || K <= 0) return 0;
  return operand_stage2_cuda(&d, &dst, &R, &K, &Kp, 1, operand_mode, st);
}

void colsum_quantize_cuda(const float* x, long long rows, int cols, long long ld, float* out, int e, int m, int rounding,
                          cudaStream_t st) {
  if (cols <= 0) return;
  CPD_DISPATCH_FORMAT(e, m, rounding, FMT, {
    colsum_quantize_kernel<decltype(FMT)><<<(cols + 31) / 32, 256, 0, st>>>(x, rows, cols, ld, out, FMT);
  });
}

}  // namespace cpd
