// lars_sm100.cu — multi-tensor LARS / momentum-SGD step on fp32 master weights,
// one persistent launch for the whole model.
//
// Reference (example/ResNet18/tools/mix.py:292-314): per tensor three torch
// norm() calls, scalar math on 0-d tensors, a momentum update, a copy_ into the
// master and a second copy_ into the model parameter — ~12 launches per tensor,
// 62 tensors per step.  Here:
//   phase A  per-chunk partial sums of w^2 and g^2 (fixed order -> deterministic)
//   ---- grid barrier ----
//   phase B  per chunk: sum its tensor's partials in chunk order (double), trust
//            ratio local_lr = coeff*|w|/(|g| + wd*|w| + eps), then
//            buf = mom*buf + lr*local_lr*(g + wd*w); w -= buf; model <- w
// With use_lars == false phase A and the barrier are skipped and the update is
// torch.optim.SGD's (optionally Nesterov).  lr is read from device memory so a
// captured CUDA graph can be replayed while the schedule advances.
#include "../api.h"
#include "../common/device_utils.cuh"

namespace cpd {

constexpr int kLarsThreads = 256;
constexpr int kVec = 4;              // float4 vectors per thread per chunk (4 x 256 x 4 = 4096 elements)

__global__ void __launch_bounds__(kLarsThreads, 2)
lars_kernel(float* master, float* model, const float* __restrict__ grad, float* mom,
            const Chunk* __restrict__ chunks, const int* __restrict__ tchunk_off, int n_chunks, int T,
            const float* __restrict__ lr_dev, float lr_host, float momentum, float wd, float coeff,
            float eps, bool use_lars, bool nesterov, const float* __restrict__ wd_mult,
            float* __restrict__ partials /*[n_chunks][2]*/, unsigned* __restrict__ bar) {
  __shared__ float red[32];
  __shared__ float s_local_lr;
  const float lr = lr_dev ? *lr_dev : lr_host;

  // LARS reads w and g twice (norms, then update).  Phase B therefore walks this CTA's chunks in REVERSE order - what
  // phase A touched last is what L2 still holds - and the very last chunk of phase A never leaves the registers.
  const int c_last = n_chunks <= (int)blockIdx.x ? -1
                                                 : (int)blockIdx.x + ((n_chunks - 1 - (int)blockIdx.x) / (int)gridDim.x) * (int)gridDim.x;
  float4 w_keep[kVec], g_keep[kVec];
  if (use_lars) {
    for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
      const Chunk ck = chunks[c];
      float w2 = 0.f, g2 = 0.f;
      {
        const int nvec = ck.len >> 2;
        const float4* wb = reinterpret_cast<const float4*>(master + ck.start);
        const float4* gb = reinterpret_cast<const float4*>(grad + ck.start);
        float4 w[kVec], g[kVec];
#pragma unroll
        for (int k = 0; k < kVec; ++k) {      // all loads first: 8 x 16 B in flight per thread
          const int v = threadIdx.x + k * kLarsThreads;
          w[k] = v < nvec ? wb[v] : make_float4(0.f, 0.f, 0.f, 0.f);
          g[k] = v < nvec ? gb[v] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < kVec; ++k) {
          w2 += w[k].x * w[k].x + w[k].y * w[k].y + w[k].z * w[k].z + w[k].w * w[k].w;
          g2 += g[k].x * g[k].x + g[k].y * g[k].y + g[k].z * g[k].z + g[k].w * g[k].w;
        }
        if (c == c_last) {
#pragma unroll
          for (int k = 0; k < kVec; ++k) { w_keep[k] = w[k]; g_keep[k] = g[k]; }
        }
      }
      w2 = block_sum_f32(w2, red);
      g2 = block_sum_f32(g2, red);
      if (threadIdx.x == 0) { partials[2 * c] = w2; partials[2 * c + 1] = g2; }
    }
    grid_barrier(bar);
  }

  for (int c = c_last; c >= 0; c -= gridDim.x) {
    const Chunk ck = chunks[c];
    const float wdt = wd * (wd_mult ? wd_mult[ck.tensor] : 1.f);
    float local_lr = 1.f;
    if (use_lars) {
      if (threadIdx.x < 32) {
        double w2 = 0.0, g2 = 0.0;
        const int c0 = tchunk_off[ck.tensor], c1 = tchunk_off[ck.tensor + 1];
        for (int k = c0 + threadIdx.x; k < c1; k += 32) { w2 += partials[2 * k]; g2 += partials[2 * k + 1]; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          w2 += __shfl_xor_sync(0xffffffffu, w2, o);
          g2 += __shfl_xor_sync(0xffffffffu, g2, o);
        }
        if (threadIdx.x == 0) {
          const float wn = (float)sqrt(w2), gn = (float)sqrt(g2);
          s_local_lr = coeff * wn / (gn + wdt * wn + eps);
        }
      }
      __syncthreads();
      local_lr = s_local_lr;
    }
    {
      const int nvec = ck.len >> 2;
      float4* wb = reinterpret_cast<float4*>(master + ck.start);
      const float4* gb = reinterpret_cast<const float4*>(grad + ck.start);
      float4* bb = reinterpret_cast<float4*>(mom + ck.start);
      float4* mb = model ? reinterpret_cast<float4*>(model + ck.start) : nullptr;
      float4 w4[kVec], g4[kVec], b4[kVec];
#pragma unroll
      for (int k = 0; k < kVec; ++k) {        // 12 x 16 B in flight per thread
        const int v = threadIdx.x + k * kLarsThreads;
        if (v < nvec) {
          b4[k] = bb[v];
          if (use_lars && c == c_last) { w4[k] = w_keep[k]; g4[k] = g_keep[k]; }
          else { w4[k] = wb[v]; g4[k] = gb[v]; }
        }
      }
#pragma unroll
      for (int k = 0; k < kVec; ++k) {
        const int v = threadIdx.x + k * kLarsThreads;
        if (v >= nvec) continue;
        float w[4] = {w4[k].x, w4[k].y, w4[k].z, w4[k].w}, g[4] = {g4[k].x, g4[k].y, g4[k].z, g4[k].w};
        float b[4] = {b4[k].x, b4[k].y, b4[k].z, b4[k].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float d = g[j] + wdt * w[j];
          float upd;
          if (use_lars) {
            b[j] = momentum * b[j] + lr * local_lr * d;
            upd = b[j];
          } else {
            b[j] = momentum * b[j] + d;
            upd = lr * (nesterov ? d + momentum * b[j] : b[j]);
          }
          w[j] = w[j] - upd;
        }
        bb[v] = make_float4(b[0], b[1], b[2], b[3]);
        wb[v] = make_float4(w[0], w[1], w[2], w[3]);
        if (mb) mb[v] = make_float4(w[0], w[1], w[2], w[3]);
      }
    }
    if (use_lars) __syncthreads();   // s_local_lr is rewritten by the next chunk
  }
}

void lars_step_cuda(float* master, float* model, const float* grad, float* mom, const Chunk* chunks,
                    const int* tchunk_off, int n_chunks, int T, const float* lr_dev, float lr,
                    float momentum, float wd, float coeff, float eps, bool use_lars, bool nesterov,
                    const float* wd_mult, float* partials, unsigned* barrier_ws, int max_ctas,
                    cudaStream_t st) {
  if (n_chunks <= 0) return;
  int grid = max_ctas > 0 ? max_ctas : device_sm_count(-1) * 2;   // 2 co-resident CTAs per SM (launch bounds), 12 x 16 B in flight per thread
  if (grid > n_chunks) grid = n_chunks;
  lars_kernel<<<grid, kLarsThreads, 0, st>>>(master, model, grad, mom, chunks, tchunk_off, n_chunks, T,
                                              lr_dev, lr, momentum, wd, coeff, eps, use_lars, nesterov,
                                              wd_mult, partials, barrier_ws);
}

}  // namespace cpd
