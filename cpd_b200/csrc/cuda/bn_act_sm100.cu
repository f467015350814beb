// bn_act_sm100.cu — training-mode BatchNorm2d fused with the residual add and the ReLU that follow it, channels_last fp32.
//
// The benchmark networks of the reference (example/ResNet18/models/resnet18_cifar.py:10-45, torchvision's ResNet-50,
// example/DavidNet/davidnet.py) are conv -> BatchNorm -> (+ shortcut) -> ReLU chains.  Run through stock modules that is
// a cuDNN BN kernel, an elementwise add and a clamp forward, and a threshold kernel plus a cuDNN BN-backward kernel
// backward: 13-16 full passes over the activation tensor per BN layer, 42 % of a ResNet-18/CIFAR step on a B200
// (profiles/launches_trip1.md: batchnorm 24 %, elementwise 18 %).  These kernels do the same arithmetic in 10-12 passes
// and 4 launches per layer:
//   bn_stats_kernel      per-channel sum / sum of squares of (x - shift)          (shift = running mean: the shifted-data
//                        formula keeps fp32 sums accurate when |mean| >> std); the LAST CTA of a channel tile reduces
//                        the per-CTA partials in a fixed order (deterministic) and writes mean, rstd, the folded
//                        scale / bias of the apply pass and the running statistics
//   bn_apply_kernel      y = relu(x * scale + bias (+ residual))
//   bn_bwd_stats_kernel  dz = dy * [y > 0];  per-channel sum dz, sum dz * xhat;  last CTA: dgamma, dbeta, coefficients
//   bn_bwd_apply_kernel  dx = gamma * rstd * (dz - mean(dz) - xhat * mean(dz * xhat))   (+ dz as the shortcut's gradient)
// x is [M = N*H*W, C] row-major (NHWC), C % 4 == 0; a warp reads 16-byte vectors along C, rows are strided over the block.
#include <cuda_runtime.h>
#include <cstdint>

#include "../api.h"
#include "../common/device_utils.cuh"

namespace cpd {

namespace {

constexpr int kBnThreads = 256;

struct BnGeom {
  int lanes;          // threads along channels (each owns 4 channels): min(C / 4, 32)
  int rows_par;       // rows processed in parallel by a block: 256 / lanes
  int ctiles;         // channel tiles of lanes * 4 channels
  int slabs;          // row slabs (grid.y)
  long long rows_per_slab;
};

inline BnGeom bn_geom(long long M, int C, int sms) {
  BnGeom g;
  g.lanes = C / 4 < 32 ? C / 4 : 32;
  g.rows_par = kBnThreads / g.lanes;
  g.ctiles = (C / 4 + g.lanes - 1) / g.lanes;
  long long want = (long long)sms * 4 / g.ctiles;
  if (want < 1) want = 1;
  long long max_slabs = (M + g.rows_par * 8 - 1) / (g.rows_par * 8);       // at least 8 rows per thread
  if (max_slabs < 1) max_slabs = 1;
  g.slabs = (int)(want < max_slabs ? want : max_slabs);
  g.rows_per_slab = (M + g.slabs - 1) / g.slabs;
  return g;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// block-level reduction of two float4 accumulators over the rows_par dimension; result valid for ty == 0
__device__ __forceinline__ void reduce_rows(float4& a, float4& b, int lanes, int rows_par, float4* sm) {
  const int lane = threadIdx.x % lanes, ty = threadIdx.x / lanes;
  sm[threadIdx.x] = a;
  sm[kBnThreads + threadIdx.x] = b;
  __syncthreads();
  if (ty == 0) {
    for (int r = 1; r < rows_par; ++r) {
      const float4 u = sm[r * lanes + lane], v = sm[kBnThreads + r * lanes + lane];
      a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
      b.x += v.x; b.y += v.y; b.z += v.z; b.w += v.w;
    }
  }
  __syncthreads();
}

// Deterministic two-level reduction of the per-slab partial sums.  The single "last CTA reduces all ~600 slabs" of the
// first version was the tail of every statistics launch: ~75 dependent-ish L2 loads per thread while 147 SMs idle
// (ncu r2t7: 23 us of SM activity inside a 42 us kernel).  Now the slabs form groups of kBnGroup; the LAST CTA of a
// group to arrive folds that group's partials (one batch of loads) into a double-precision group partial, and the last
// group to finish folds the <= 37 group partials.  Every sum runs in a fixed order (slab order inside a group, group
// order across groups), whichever CTA happens to execute it: bitwise reproducible.
constexpr int kBnGroup = 16;

__device__ __forceinline__ double4 ldcg_d4(const double* p) {
  const double2 a = __ldcg(reinterpret_cast<const double2*>(p)), b = __ldcg(reinterpret_cast<const double2*>(p) + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ double4 ldcg_as_d4(const float* p) {
  const float4 v = __ldcg(reinterpret_cast<const float4*>(p));
  return make_double4(v.x, v.y, v.z, v.w);
}
__device__ __forceinline__ double4 ldcg_as_d4(const double* p) { return ldcg_d4(p); }

// entries [first, first + count) of src[.][2][C] (float or double) -> per-channel sums in a / b of the g == 0 threads:
// thread (lane, g) adds entries g, g + G, ... (4 in flight), the G partial sums are then combined in g order
template <typename T>
__device__ __forceinline__ void reduce_entries(const T* __restrict__ src, int first, int count, int C, int lanes, double4& a, double4& b,
                                               double4* scratch) {
  const int G = kBnThreads / lanes;
  const int lane = threadIdx.x % lanes, g = threadIdx.x / lanes;
  const int c0 = (blockIdx.x * lanes + lane) * 4;
  a = make_double4(0.0, 0.0, 0.0, 0.0); b = a;
  if (c0 < C) {
    int e = g;
    for (; e + 3 * G < count; e += 4 * G) {
      double4 u[4], v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        u[j] = ldcg_as_d4(src + ((long long)(first + e + j * G) * 2 + 0) * C + c0);     // other SMs wrote these
        v[j] = ldcg_as_d4(src + ((long long)(first + e + j * G) * 2 + 1) * C + c0);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a.x += u[j].x; a.y += u[j].y; a.z += u[j].z; a.w += u[j].w;
        b.x += v[j].x; b.y += v[j].y; b.z += v[j].z; b.w += v[j].w;
      }
    }
    for (; e < count; e += G) {
      const double4 u = ldcg_as_d4(src + ((long long)(first + e) * 2 + 0) * C + c0);
      const double4 v = ldcg_as_d4(src + ((long long)(first + e) * 2 + 1) * C + c0);
      a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
      b.x += v.x; b.y += v.y; b.z += v.z; b.w += v.w;
    }
  }
  scratch[threadIdx.x] = a;
  scratch[kBnThreads + threadIdx.x] = b;
  __syncthreads();
  if (g == 0 && c0 < C) {
    for (int j = 1; j < G; ++j) {
      const double4 u = scratch[j * lanes + lane], v = scratch[kBnThreads + j * lanes + lane];
      a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
      b.x += v.x; b.y += v.y; b.z += v.z; b.w += v.w;
    }
  }
  __syncthreads();
}

// Called by every CTA after it wrote part[blockIdx.y].  Returns true in exactly one CTA per channel tile - the one that
// finishes the reduction - with the per-channel totals in s_sum / s_sq (shared, double).
// counters (per channel tile): [0] groups finished, [1 + grp] slabs of group grp arrived; all return to 0.
__device__ __forceinline__ bool arrive_and_reduce(const float* __restrict__ part, double* __restrict__ gpart, unsigned* __restrict__ counters,
                                                  int C, int lanes, double* s_sum, double* s_sq, double4* scratch) {
  __shared__ bool last;
  const int slabs = (int)gridDim.y, groups = (slabs + kBnGroup - 1) / kBnGroup;
  const int grp = (int)blockIdx.y / kBnGroup;
  const int gfirst = grp * kBnGroup, gcount = min(kBnGroup, slabs - gfirst);
  unsigned* cnt = counters + (long long)blockIdx.x * (1 + groups);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(&cnt[1 + grp], 1u);
    last = done == (unsigned)gcount - 1u;
    if (last) cnt[1 + grp] = 0u;                                // ready for the next launch / graph replay
  }
  __syncthreads();
  if (!last) return false;
  __threadfence();
  const int lane = threadIdx.x % lanes, g = threadIdx.x / lanes;
  const int c0 = (blockIdx.x * lanes + lane) * 4;
  double4 a, b;
  reduce_entries(part, gfirst, gcount, C, lanes, a, b, scratch);
  if (groups > 1) {
    if (g == 0 && c0 < C) {
      double* o = gpart + ((long long)grp * 2) * C + c0;
      *reinterpret_cast<double2*>(o) = make_double2(a.x, a.y); *reinterpret_cast<double2*>(o + 2) = make_double2(a.z, a.w);
      *reinterpret_cast<double2*>(o + C) = make_double2(b.x, b.y); *reinterpret_cast<double2*>(o + C + 2) = make_double2(b.z, b.w);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned done = atomicAdd(&cnt[0], 1u);
      last = done == (unsigned)groups - 1u;
      if (last) cnt[0] = 0u;
    }
    __syncthreads();
    if (!last) return false;
    __threadfence();
    reduce_entries(gpart, 0, groups, C, lanes, a, b, scratch);
  }
  if (g == 0 && c0 < C) {
    s_sum[lane * 4 + 0] = a.x; s_sum[lane * 4 + 1] = a.y; s_sum[lane * 4 + 2] = a.z; s_sum[lane * 4 + 3] = a.w;
    s_sq[lane * 4 + 0] = b.x; s_sq[lane * 4 + 1] = b.y; s_sq[lane * 4 + 2] = b.z; s_sq[lane * 4 + 3] = b.w;
  }
  __syncthreads();
  return true;
}

// ---- forward: statistics ------------------------------------------------------------------------------------------
// part: [slabs][2][C]; stats out: mean[C], rstd[C], scale[C] = gamma * rstd, bias[C] = beta - mean * scale
__global__ void __launch_bounds__(kBnThreads)
bn_stats_kernel(const float* __restrict__ x, long long M, int C, long long rows_per_slab, int lanes,
                const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ running_mean,
                float* __restrict__ running_var, float momentum, float eps, float* __restrict__ part,
                double* __restrict__ gpart, unsigned* __restrict__ counters, float* __restrict__ mean, float* __restrict__ rstd,
                float* __restrict__ scale, float* __restrict__ bias, long long* __restrict__ num_batches_tracked) {
  __shared__ float4 sm[2 * kBnThreads];
  const int rows_par = kBnThreads / lanes;
  const int lane = threadIdx.x % lanes, ty = threadIdx.x / lanes;
  const int c0 = (blockIdx.x * lanes + lane) * 4;
  const bool cok = c0 < C;
  const float4 sh = cok ? ld4(running_mean + c0) : make_float4(0.f, 0.f, 0.f, 0.f);       // shift of the shifted-data sums
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  // Rows are dealt to the grid ROUND-ROBIN in groups of rows_par (not one contiguous slab per CTA): at any moment the
  // whole grid reads one compact window of the tensor that moves forward - the access pattern of the apply kernels,
  // which stream at 6.2 TB/s where per-CTA slabs (592 concurrent streams) reached 3.2 TB/s (ncu r2t4_prof_bn_stats_kernel).
  (void)rows_per_slab;
  const long long step = (long long)gridDim.y * rows_par;
  if (cok) {
    long long r = (long long)blockIdx.y * rows_par + ty;
    // 8 independent 16-byte loads in flight per thread
    for (; r + 7LL * step < M; r += 8LL * step) {
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = ld4(x + (r + (long long)j * step) * C + c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float a = v[j].x - sh.x, b = v[j].y - sh.y, c = v[j].z - sh.z, d = v[j].w - sh.w;
        s.x += a; s.y += b; s.z += c; s.w += d;
        q.x += a * a; q.y += b * b; q.z += c * c; q.w += d * d;
      }
    }
    for (; r < M; r += step) {
      const float4 v = ld4(x + r * C + c0);
      const float a = v.x - sh.x, b = v.y - sh.y, c = v.z - sh.z, d = v.w - sh.w;
      s.x += a; s.y += b; s.z += c; s.w += d;
      q.x += a * a; q.y += b * b; q.z += c * c; q.w += d * d;
    }
  }
  reduce_rows(s, q, lanes, rows_par, sm);
  if (ty == 0 && cok) {
    *reinterpret_cast<float4*>(part + ((long long)blockIdx.y * 2 + 0) * C + c0) = s;
    *reinterpret_cast<float4*>(part + ((long long)blockIdx.y * 2 + 1) * C + c0) = q;
  }
  // fixed-order two-level reduction of the slab partials; one CTA per channel tile comes out with the totals
  __shared__ double s_sum[128], s_sq[128];
  __shared__ double4 scratch[2 * kBnThreads];
  if (!arrive_and_reduce(part, gpart, counters, C, lanes, s_sum, s_sq, scratch)) return;
  if (num_batches_tracked && blockIdx.x == 0 && threadIdx.x == 0) *num_batches_tracked += 1;     // nn.BatchNorm2d's counter
  const int nch = lanes * 4;
  for (int i = threadIdx.x; i < nch; i += kBnThreads) {
    const int c = blockIdx.x * nch + i;
    if (c >= C) continue;
    const double ds = s_sum[i], dq = s_sq[i];
    const double shift = (double)running_mean[c];
    const double ms = ds / (double)M;                             // mean of (x - shift)
    double var = dq / (double)M - ms * ms;                        // biased variance
    if (var < 0.0) var = 0.0;
    const float mu = (float)(ms + shift);
    const float rs = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * rs;
    mean[c] = mu; rstd[c] = rs; scale[c] = sc; bias[c] = beta[c] - mu * sc;
    const double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mu;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// ---- forward: apply -----------------------------------------------------------------------------------------------
template <bool RELU, bool RES>
__global__ void __launch_bounds__(kBnThreads)
bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ res, float* __restrict__ y, long long n4, int C4,
                const float* __restrict__ scale, const float* __restrict__ bias, unsigned* __restrict__ mask) {
  const long long stride = (long long)gridDim.x * kBnThreads;
  // ReLU bitmask for the backward pass (1 bit per element instead of re-reading y: 32x less traffic there): 8 consecutive
  // threads hold 8 consecutive vectors = 32 elements = one word.  n4 % 8 == 0 (checked on the host), so the 8 lanes of a
  // group are all in range or all out of range and the group-wide shuffles are well defined.
  const unsigned gmask = 0xFFu << (threadIdx.x & 24);
  // two independent vectors per thread per iteration: 32-64 B in flight per thread keeps HBM busy at 2048 threads / SM
  for (long long i = (long long)blockIdx.x * kBnThreads + threadIdx.x; i < n4; i += 2 * stride) {
    const long long j = i + stride;
    const bool two = j < n4;
    const float4 v0 = ld4(x + i * 4);
    const float4 v1 = two ? ld4(x + j * 4) : v0;
    float4 r0, r1;
    if (RES) { r0 = ld4(res + i * 4); r1 = two ? ld4(res + j * 4) : r0; }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 1 && !two) break;
      const long long e = h ? j : i;
      const float4 v = h ? v1 : v0;
      const int c0 = (int)(e % C4) * 4;
      const float4 sc = ld4(scale + c0), bi = ld4(bias + c0);
      float4 o = make_float4(v.x * sc.x + bi.x, v.y * sc.y + bi.y, v.z * sc.z + bi.z, v.w * sc.w + bi.w);
      if (RES) {
        const float4 r = h ? r1 : r0;
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      if (RELU) {
        if (mask) {
          unsigned nib = (o.x > 0.f ? 1u : 0u) | (o.y > 0.f ? 2u : 0u) | (o.z > 0.f ? 4u : 0u) | (o.w > 0.f ? 8u : 0u);
          nib <<= 4 * (threadIdx.x & 7);
          nib |= __shfl_xor_sync(gmask, nib, 1);
          nib |= __shfl_xor_sync(gmask, nib, 2);
          nib |= __shfl_xor_sync(gmask, nib, 4);
          if ((threadIdx.x & 7) == 0) mask[e >> 3] = nib;
        }
        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
      }
      *reinterpret_cast<float4*>(y + e * 4) = o;
    }
  }
}

// ---- backward: statistics -----------------------------------------------------------------------------------------
// coef out: c1[C] = mean(dz), c2[C] = mean(dz * xhat); dgamma[C], dbeta[C]
template <bool RELU>
__global__ void __launch_bounds__(kBnThreads)
bn_bwd_stats_kernel(const float* __restrict__ dy, const float* __restrict__ y, const unsigned* __restrict__ mask,
                    const float* __restrict__ x, long long M, int C,
                    long long rows_per_slab, int lanes, const float* __restrict__ mean, const float* __restrict__ rstd,
                    float* __restrict__ part, double* __restrict__ gpart, unsigned* __restrict__ counters, float* __restrict__ dgamma,
                    float* __restrict__ dbeta, float* __restrict__ c1, float* __restrict__ c2) {
  __shared__ float4 sm[2 * kBnThreads];
  const int rows_par = kBnThreads / lanes;
  const int lane = threadIdx.x % lanes, ty = threadIdx.x / lanes;
  const int c0 = (blockIdx.x * lanes + lane) * 4;
  const bool cok = c0 < C;
  const float4 mu = cok ? ld4(mean + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 rs = cok ? ld4(rstd + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  (void)rows_per_slab;
  const long long step = (long long)gridDim.y * rows_par;      // rows dealt round-robin to the grid (see bn_stats_kernel)
  if (cok) {
    // 4 rows x (dy, x, mask | y) = up to 12 independent loads in flight per thread
    for (long long r = (long long)blockIdx.y * rows_par + ty; r < M; r += 4LL * step) {
      float4 g[4], xv[4], yv[4];
      unsigned mb[4];
      bool ok[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long rj = r + (long long)j * step;
        ok[j] = rj < M;
        const long long a = (ok[j] ? rj : r) * C + c0;
        g[j] = ld4(dy + a);
        xv[j] = ld4(x + a);
        if (RELU) {
          if (mask) mb[j] = (mask[a >> 5] >> (4 * (int)((a >> 2) & 7))) & 0xFu;       // vector a/4: word (a/4)/8, nibble (a/4)%8
          else yv[j] = ld4(y + a);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (!ok[j]) continue;
        if (RELU) {
          if (mask) {
            g[j].x = (mb[j] & 1u) ? g[j].x : 0.f; g[j].y = (mb[j] & 2u) ? g[j].y : 0.f;
            g[j].z = (mb[j] & 4u) ? g[j].z : 0.f; g[j].w = (mb[j] & 8u) ? g[j].w : 0.f;
          } else {
            g[j].x = yv[j].x > 0.f ? g[j].x : 0.f; g[j].y = yv[j].y > 0.f ? g[j].y : 0.f;
            g[j].z = yv[j].z > 0.f ? g[j].z : 0.f; g[j].w = yv[j].w > 0.f ? g[j].w : 0.f;
          }
        }
        s.x += g[j].x; s.y += g[j].y; s.z += g[j].z; s.w += g[j].w;
        q.x += g[j].x * (xv[j].x - mu.x) * rs.x; q.y += g[j].y * (xv[j].y - mu.y) * rs.y;
        q.z += g[j].z * (xv[j].z - mu.z) * rs.z; q.w += g[j].w * (xv[j].w - mu.w) * rs.w;
      }
    }
  }
  reduce_rows(s, q, lanes, rows_par, sm);
  if (ty == 0 && cok) {
    *reinterpret_cast<float4*>(part + ((long long)blockIdx.y * 2 + 0) * C + c0) = s;
    *reinterpret_cast<float4*>(part + ((long long)blockIdx.y * 2 + 1) * C + c0) = q;
  }
  __shared__ double s_sum[128], s_sq[128];
  __shared__ double4 scratch[2 * kBnThreads];
  if (!arrive_and_reduce(part, gpart, counters, C, lanes, s_sum, s_sq, scratch)) return;
  const int nch = lanes * 4;
  for (int i = threadIdx.x; i < nch; i += kBnThreads) {
    const int c = blockIdx.x * nch + i;
    if (c >= C) continue;
    const double ds = s_sum[i], dq = s_sq[i];
    dbeta[c] = (float)ds;
    dgamma[c] = (float)dq;
    c1[c] = (float)(ds / (double)M);
    c2[c] = (float)(dq / (double)M);
  }
}

// ---- backward: apply ----------------------------------------------------------------------------------------------
template <bool RELU, bool RES>
__global__ void __launch_bounds__(kBnThreads)
bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ y, const unsigned* __restrict__ mask,
                    const float* __restrict__ x, float* __restrict__ dx,
                    float* __restrict__ dres, long long n4, int C4, const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ gamma, const float* __restrict__ c1, const float* __restrict__ c2) {
  const long long stride = (long long)gridDim.x * kBnThreads;
  for (long long i = (long long)blockIdx.x * kBnThreads + threadIdx.x; i < n4; i += stride) {
    const int c0 = (int)(i % C4) * 4;
    float4 g = ld4(dy + i * 4);
    const float4 v = ld4(x + i * 4);
    if (RELU) {
      if (mask) {
        const unsigned mb = (mask[i >> 3] >> (4 * (int)(i & 7))) & 0xFu;
        g.x = (mb & 1u) ? g.x : 0.f; g.y = (mb & 2u) ? g.y : 0.f; g.z = (mb & 4u) ? g.z : 0.f; g.w = (mb & 8u) ? g.w : 0.f;
      } else {
        const float4 o = ld4(y + i * 4);
        g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f; g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
      }
    }
    if (RES) *reinterpret_cast<float4*>(dres + i * 4) = g;
    const float4 mu = ld4(mean + c0), rs = ld4(rstd + c0), ga = ld4(gamma + c0), a = ld4(c1 + c0), b = ld4(c2 + c0);
    float4 o;
    o.x = ga.x * rs.x * (g.x - a.x - (v.x - mu.x) * rs.x * b.x);
    o.y = ga.y * rs.y * (g.y - a.y - (v.y - mu.y) * rs.y * b.y);
    o.z = ga.z * rs.z * (g.z - a.z - (v.z - mu.z) * rs.z * b.z);
    o.w = ga.w * rs.w * (g.w - a.w - (v.w - mu.w) * rs.w * b.w);
    *reinterpret_cast<float4*>(dx + i * 4) = o;
  }
}

int ew_grid(long long n4, int sms) {
  long long b = (n4 + kBnThreads - 1) / kBnThreads;
  const long long cap = (long long)sms * 8;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace

// workspace: unsigned counters[kBnCounters] | double gpart[groups * 2 * C] | float part[slabs * 2 * C].  The counters sit at a
// FIXED offset: they are the only part that must be zero between launches (the kernels re-arm them), so one zero-
// initialised buffer serves any batch size - the last, smaller batch of an epoch uses a different slab geometry.
// stats: float[4 * C] = mean | rstd | scale | bias
constexpr int kBnCounters = 1024;
struct BnWs { float* part; double* gpart; unsigned* counters; };
inline int bn_ws_groups(const BnGeom& g) { return (g.slabs + kBnGroup - 1) / kBnGroup; }
inline BnWs bn_ws(float* workspace, const BnGeom& g, int C) {
  BnWs w;
  w.counters = reinterpret_cast<unsigned*>(workspace);
  w.gpart = reinterpret_cast<double*>(workspace + kBnCounters);
  w.part = reinterpret_cast<float*>(w.gpart + (long long)bn_ws_groups(g) * 2 * C);
  return w;
}
inline bool bn_ws_ok(const float* workspace, const BnGeom& g) {
  return !(reinterpret_cast<uintptr_t>(workspace) & 7u) && (long long)g.ctiles * (1 + bn_ws_groups(g)) <= kBnCounters;
}
long long bn_workspace_floats(long long M, int C) {
  const BnGeom g = bn_geom(M, C, device_sm_count(-1));
  return kBnCounters + 2LL * bn_ws_groups(g) * 2 * C + (long long)g.slabs * 2 * C + 8;
}

int bn_act_forward_cuda(const float* x, const float* res, float* y, long long M, int C, const float* gamma, const float* beta,
                        float* running_mean, float* running_var, long long* num_batches_tracked, float momentum, float eps,
                        bool relu, float* stats, float* workspace, unsigned* relu_mask, cudaStream_t st) {
  if (M <= 0 || C <= 0) return 0;
  if (C % 4) return -1;
  const int sms = device_sm_count(-1);
  const BnGeom g = bn_geom(M, C, sms);
  if (!bn_ws_ok(workspace, g)) return -3;
  const BnWs w = bn_ws(workspace, g, C);
  float* part = w.part;
  unsigned* counters = w.counters;
  float *mean = stats, *rstd = stats + C, *scale = stats + 2 * C, *bias = stats + 3 * C;
  bn_stats_kernel<<<dim3(g.ctiles, g.slabs), kBnThreads, 0, st>>>(x, M, C, g.rows_per_slab, g.lanes, gamma, beta, running_mean, running_var,
                                                                   momentum, eps, part, w.gpart, counters, mean, rstd, scale, bias,
                                                                   num_batches_tracked);
  const long long n4 = M * C / 4;
  if (relu_mask && (n4 % 8)) return -2;                  // the bitmask path packs 8 vectors per word
  const int grid = ew_grid(n4, sms);
  if (relu && res) bn_apply_kernel<true, true><<<grid, kBnThreads, 0, st>>>(x, res, y, n4, C / 4, scale, bias, relu_mask);
  else if (relu) bn_apply_kernel<true, false><<<grid, kBnThreads, 0, st>>>(x, res, y, n4, C / 4, scale, bias, relu_mask);
  else if (res) bn_apply_kernel<false, true><<<grid, kBnThreads, 0, st>>>(x, res, y, n4, C / 4, scale, bias, nullptr);
  else bn_apply_kernel<false, false><<<grid, kBnThreads, 0, st>>>(x, res, y, n4, C / 4, scale, bias, nullptr);
  return 0;
}

// grads: float[4 * C] = dgamma | dbeta | c1 | c2
int bn_act_backward_cuda(const float* dy, const float* y, const unsigned* relu_mask, const float* x, float* dx, float* dres, long long M,
                         int C, const float* gamma, const float* stats, bool relu, float* grads, float* workspace, cudaStream_t st) {
  if (M <= 0 || C <= 0) return 0;
  if (C % 4) return -1;
  const int sms = device_sm_count(-1);
  const BnGeom g = bn_geom(M, C, sms);
  if (!bn_ws_ok(workspace, g)) return -3;
  const BnWs w = bn_ws(workspace, g, C);
  float* part = w.part;
  unsigned* counters = w.counters;
  const float *mean = stats, *rstd = stats + C;
  float *dgamma = grads, *dbeta = grads + C, *c1 = grads + 2 * C, *c2 = grads + 3 * C;
  if (relu)
    bn_bwd_stats_kernel<true><<<dim3(g.ctiles, g.slabs), kBnThreads, 0, st>>>(dy, y, relu_mask, x, M, C, g.rows_per_slab, g.lanes, mean, rstd, part,
                                                                                w.gpart, counters, dgamma, dbeta, c1, c2);
  else
    bn_bwd_stats_kernel<false><<<dim3(g.ctiles, g.slabs), kBnThreads, 0, st>>>(dy, y, nullptr, x, M, C, g.rows_per_slab, g.lanes, mean, rstd, part,
                                                                                 w.gpart, counters, dgamma, dbeta, c1, c2);
  const long long n4 = M * C / 4;
  const int grid = ew_grid(n4, sms);
  if (relu && dres) bn_bwd_apply_kernel<true, true><<<grid, kBnThreads, 0, st>>>(dy, y, relu_mask, x, dx, dres, n4, C / 4, mean, rstd, gamma, c1, c2);
  else if (relu) bn_bwd_apply_kernel<true, false><<<grid, kBnThreads, 0, st>>>(dy, y, relu_mask, x, dx, dres, n4, C / 4, mean, rstd, gamma, c1, c2);
  else if (dres) bn_bwd_apply_kernel<false, true><<<grid, kBnThreads, 0, st>>>(dy, y, nullptr, x, dx, dres, n4, C / 4, mean, rstd, gamma, c1, c2);
  else bn_bwd_apply_kernel<false, false><<<grid, kBnThreads, 0, st>>>(dy, y, nullptr, x, dx, dres, n4, C / 4, mean, rstd, gamma, c1, c2);
  return 0;
}

}  // namespace cpd
