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

__global__ void __launch_bounds__(kLarsThreads)
lars_kernel(float* master, float* model, const float* __restrict__ grad, float* mom,
            const Chunk* __restrict__ chunks, const int* __restrict__ tchunk_off, int n_chunks, int T,
            const float* __restrict__ lr_dev, float lr_host, float momentum, float wd, float coeff,
            float eps, bool use_lars, bool nesterov, const float* __restrict__ wd_mult,
            float* __restrict__ partials /*[n_chunks][2]*/, unsigned* __restrict__ bar) {
  __shared__ float red[32];
  __shared__ float s_local_lr;
  const float lr = lr_dev ? *lr_dev : lr_host;

  if (use_lars) {
    for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
      const Chunk ck = chunks[c];
      float w2 = 0.f, g2 = 0.f;
      for (int i = threadIdx.x * 4; i < ck.len; i += kLarsThreads * 4) {
        const float4 w = *reinterpret_cast<const float4*>(master + ck.start + i);
        const float4 g = *reinterpret_cast<const float4*>(grad + ck.start + i);
        w2 += w.x * w.x + w.y * w.y + w.z * w.z + w.w * w.w;
        g2 += g.x * g.x + g.y * g.y + g.z * g.z + g.w * g.w;
      }
      w2 = block_sum_f32(w2, red);
      g2 = block_sum_f32(g2, red);
      if (threadIdx.x == 0) { partials[2 * c] = w2; partials[2 * c + 1] = g2; }
    }
    grid_barrier(bar);
  }

  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
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
    for (int i = threadIdx.x * 4; i < ck.len; i += kLarsThreads * 4) {
      const long long o = ck.start + i;
      const float4 w4 = *reinterpret_cast<const float4*>(master + o);
      const float4 g4 = *reinterpret_cast<const float4*>(grad + o);
      const float4 b4 = *reinterpret_cast<const float4*>(mom + o);
      float w[4] = {w4.x, w4.y, w4.z, w4.w}, g[4] = {g4.x, g4.y, g4.z, g4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
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
      *reinterpret_cast<float4*>(mom + o) = make_float4(b[0], b[1], b[2], b[3]);
      *reinterpret_cast<float4*>(master + o) = make_float4(w[0], w[1], w[2], w[3]);
      if (model) *reinterpret_cast<float4*>(model + o) = make_float4(w[0], w[1], w[2], w[3]);
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
  int grid = max_ctas > 0 ? max_ctas : device_sm_count(-1) * 2;
  if (grid > n_chunks) grid = n_chunks;
  lars_kernel<<<grid, kLarsThreads, 0, st>>>(master, model, grad, mom, chunks, tchunk_off, n_chunks, T,
                                              lr_dev, lr, momentum, wd, coeff, eps, use_lars, nesterov,
                                              wd_mult, partials, barrier_ws);
}

}  // namespace cpd
