// cpu_ops.cpp — C++ CPU implementations of the CPD numerics (quantize, exact and
// tile-emulated GEMM, APS + ordered/Kahan fold, LARS).  They are (a) the CPU
// backend of the public API (the reference has none: quant_function.py:24-30
// raises for CPU tensors) and (b) the oracle the CUDA kernels are tested against.
#include <algorithm>
#include <cmath>
#include <thread>
#include <vector>

#define CPD_NO_CUDA
#include "../api.h"
#include "../common/fpformat.h"

namespace cpd {

namespace {
// Minimal fork-join helper (the image ships no OpenMP runtime to link against).
template <class Fn>
void parallel_for(int64_t n, int64_t grain, Fn fn) {
  int64_t nt = (int64_t)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if (nt > 16) nt = 16;
  if (n < 2 * grain || nt == 1) { fn(0, n); return; }
  nt = std::min<int64_t>(nt, (n + grain - 1) / grain);
  const int64_t per = (n + nt - 1) / nt;
  std::vector<std::thread> th;
  for (int64_t t = 1; t < nt; ++t) {
    const int64_t b = t * per, e = std::min(n, b + per);
    if (b < e) th.emplace_back([=] { fn(b, e); });
  }
  fn(0, std::min(n, per));
  for (auto& x : th) x.join();
}
}  // namespace

void quantize_cpu(const float* in, float* out, int64_t n, int e, int m, int rounding) {
  const Format f = make_format(e, m, rounding);
  parallel_for(n, 1 << 15, [=](int64_t b, int64_t e_) {
    for (int64_t i = b; i < e_; ++i) out[i] = quantize(in[i], f);
  });
}

void quantize_scaled_cpu(const float* in, float* out, int64_t n, int e, int m, int rounding, int shift) {
  const Format f = make_format(e, m, rounding);
  const float sc = pow2i(shift);
  parallel_for(n, 1 << 15, [=](int64_t b, int64_t e_) {
    for (int64_t i = b; i < e_; ++i) out[i] = quantize(in[i] * sc, f);
  });
}

// Reference semantics (float_kernel.cu:181-195): for k = 0..K-1 sequentially
//   p = Q(a*b); y = Q(p - c); t = Q(s + y); c = Q(Q(t - s) - y); s = t
// `kahan=false` gives the plain variant s = Q(s + Q(a*b)) that the reference
// keeps commented out (float_kernel.cu:196-200).
void gemm_exact_cpu(const float* A, int64_t sa0, int64_t sa1, const float* B, int64_t sb0, int64_t sb1,
                    float* C, int64_t ldc, int M, int N, int K, int e, int m, int rounding, bool kahan) {
  const Format f = make_format(e, m, rounding);
  FmtR q{f};
  const int64_t grain = std::max<int64_t>(1, (1 << 14) / std::max<int64_t>(1, (int64_t)N * K));
  parallel_for(M, grain, [=](int64_t r0, int64_t r1) {
  for (int64_t i = r0; i < r1; ++i) {
    std::vector<float> s(N, 0.f), c(N, 0.f);
    for (int k = 0; k < K; ++k) {
      const float a = A[i * sa0 + k * sa1];
      const float* brow = B + (int64_t)k * sb0;
      // the product is rounded to fp32 first (built with -ffp-contract=off), then cast
      if (kahan) {
        for (int j = 0; j < N; ++j) fold_kahan(q, s[j], c[j], q.q(a * brow[j * sb1]));
      } else {
        for (int j = 0; j < N; ++j) fold_plain(q, s[j], q.q(a * brow[j * sb1]));
      }
    }
    for (int j = 0; j < N; ++j) C[i * ldc + j] = s[j];
  }
  });
}

void gemm_tile_cpu(const float* A, int64_t sa0, int64_t sa1, const float* B, int64_t sb0, int64_t sb1,
                   float* C, int64_t ldc, int M, int N, int K, int e, int m, int rounding, bool kahan,
                   int k_slice, int op_e, int op_m, int splits) {
  const Format f = make_format(e, m, rounding);
  const Format fo = make_format(op_e, op_m, ROUND_IEEE);
  const bool round_ops = !(op_e == 8 && op_m == 23);
  FmtR q{f};
  if (k_slice <= 0) k_slice = K;
  // split-K as the tensor-core launcher partitions it: whole K-slices per split, every split accumulates from zero,
  // then the splits are folded IN ORDER with the same cast (C = fold_z Q(C (+) part_z))
  const int slices = (K + k_slice - 1) / k_slice;
  if (splits < 1) splits = 1;
  if (splits > slices) splits = slices;
  const int slices_per_split = (slices + splits - 1) / splits;
  const int64_t k_per_split = (int64_t)slices_per_split * k_slice;
  splits = (int)((K + k_per_split - 1) / k_per_split);
  const int64_t grain = std::max<int64_t>(1, (1 << 14) / std::max<int64_t>(1, (int64_t)N * K));
  parallel_for(M, grain, [=](int64_t r0, int64_t r1) {
  for (int64_t i = r0; i < r1; ++i) {
    std::vector<float> tot(N, 0.f), totc(N, 0.f);
    std::vector<double> part(N);
    for (int z = 0; z < splits; ++z) {
      std::vector<float> s(N, 0.f), c(N, 0.f);
      const int kz0 = (int)(z * k_per_split), kz1 = (int)std::min<int64_t>(K, (z + 1) * k_per_split);
      for (int k0 = kz0; k0 < kz1; k0 += k_slice) {
        const int k1 = std::min(kz1, k0 + k_slice);
        std::fill(part.begin(), part.end(), 0.0);
        for (int k = k0; k < k1; ++k) {
          float a = A[i * sa0 + k * sa1];
          if (round_ops) a = quantize(a, fo);
          const float* brow = B + (int64_t)k * sb0;
          for (int j = 0; j < N; ++j) {
            float b = brow[j * sb1];
            if (round_ops) b = quantize(b, fo);
            part[j] += (double)a * (double)b;    // stands in for the tensor core's wide accumulate
          }
        }
        for (int j = 0; j < N; ++j) {
          const float p = q.q((float)part[j]);
          if (kahan) fold_kahan(q, s[j], c[j], p);
          else fold_plain(q, s[j], p);
        }
      }
      if (splits == 1) { tot = s; break; }
      for (int j = 0; j < N; ++j) {
        if (kahan) fold_kahan(q, tot[j], totc[j], s[j]);
        else fold_plain(q, tot[j], s[j]);
      }
    }
    for (int j = 0; j < N; ++j) C[i * ldc + j] = tot[j];
  }
  });
}

// SURVEY §2.9.3 / dist_util.py:22-89.  srcs[r] is rank r's flat gradient arena.
void aps_fold_cpu(const float* const* srcs, int W, float* out, const int64_t* seg_off, int T, int e,
                  int m, int rounding, bool use_aps, bool kahan, float world_mult, int* shifts_out) {
  const Format f = make_format(e, m, rounding);
  FmtR q{f};
  const bool ident = is_identity(e, m) && !kahan;
  for (int t = 0; t < T; ++t) {
    const int64_t b = seg_off[t], en = seg_off[t + 1];
    int shift = 0;
    if (use_aps) {
      float me = -INFINITY;
      for (int r = 0; r < W; ++r) {
        float amax = 0.f;
        bool nan = false;
        for (int64_t i = b; i < en; ++i) {
          const float v = std::fabs(srcs[r][i]);
          if (v != v) nan = true;
          amax = v > amax ? v : amax;
        }
        const float mr = nan ? NAN : aps_local_exponent(amax, world_mult);
        if (mr != mr) me = NAN;
        else if (!(me != me) && mr > me) me = mr;
      }
      shift = aps_shift_from_exponent(me, e);
    }
    if (shifts_out) shifts_out[t] = shift;
    const float sc = pow2i(shift), inv = pow2i(-shift);
    parallel_for(en - b, 1 << 14, [=](int64_t i0, int64_t i1) {
    for (int64_t i = b + i0; i < b + i1; ++i) {
      float s = 0.f, c = 0.f;
      for (int r = 0; r < W; ++r) {
        float g = srcs[r][i];
        if (use_aps) g = q.q(g * sc);
        if (ident) s = s + g;
        else if (kahan) fold_kahan(q, s, c, g);
        else fold_plain(q, s, g);
      }
      out[i] = use_aps ? s * inv : s;
    }
    });
  }
}

// mix.py:297-314:  local_lr = coeff*|w|/(|g| + wd*|w|);  buf = mom*buf + lr*local_lr*(g + wd*w);
// w -= buf; model <- master.  eps (default 0) guards the zero-norm case the reference NaNs on.
// With use_lars=false this is torch.optim.SGD (optionally Nesterov) on the masters:
//   d = g + wd*w; buf = mom*buf + d; w -= lr*(nesterov ? d + mom*buf : buf)
void lars_step_cpu(float* master, float* model, const float* grad, float* mom, const int64_t* seg_off,
                   int T, float lr, float momentum, float wd, float coeff, float eps, bool use_lars,
                   bool nesterov, const float* wd_mult) {
  for (int t = 0; t < T; ++t) {
    const int64_t b = seg_off[t], en = seg_off[t + 1];
    const float wdt = wd * (wd_mult ? wd_mult[t] : 1.f);
    float local_lr = 1.f;
    if (use_lars) {
      double w2 = 0, g2 = 0;
      for (int64_t i = b; i < en; ++i) { w2 += (double)master[i] * master[i]; g2 += (double)grad[i] * grad[i]; }
      const float wn = (float)std::sqrt(w2), gn = (float)std::sqrt(g2);
      local_lr = coeff * wn / (gn + wdt * wn + eps);
    }
    for (int64_t i = b; i < en; ++i) {
      const float w = master[i];
      const float d = grad[i] + wdt * w;
      float upd;
      if (use_lars) {
        mom[i] = momentum * mom[i] + lr * local_lr * d;
        upd = mom[i];
      } else {
        mom[i] = momentum * mom[i] + d;
        upd = lr * (nesterov ? d + momentum * mom[i] : mom[i]);
      }
      master[i] = w - upd;
      if (model) model[i] = master[i];
    }
  }
}

}  // namespace cpd
