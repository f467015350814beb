// Self-test of the C++ CPU numerics for the sanitizer builds (tools/sanitize_cpu.sh: ASAN + UBSAN, and
// -Wall -Wextra -Werror).  It exercises every exported CPU op on small, ragged and adversarial inputs
// (specials, fp32 denormals, zero-size, strided operands) and checks format-independent invariants, so a
// memory error or undefined shift in the format core (cf. the reference's `1 << -1`, SURVEY §2.9.1
// quirk C) fails the run even where the python tests would only see a plausible number.
#define CPD_NO_CUDA
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <random>
#include <vector>

#include "../../cpd_b200/csrc/api.h"
#include "../../cpd_b200/csrc/common/fpformat.h"

static int failures = 0;
#define CHECK(cond, ...)                                                     \
  do {                                                                       \
    if (!(cond)) {                                                           \
      ++failures;                                                            \
      std::fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__);              \
      std::fprintf(stderr, __VA_ARGS__);                                     \
      std::fprintf(stderr, "\n");                                            \
    }                                                                        \
  } while (0)

static uint32_t bits(float x) { uint32_t u; std::memcpy(&u, &x, 4); return u; }
static float from_bits(uint32_t u) { float x; std::memcpy(&x, &u, 4); return x; }
static bool same(float a, float b) { return bits(a) == bits(b) || (std::isnan(a) && std::isnan(b)); }

static std::vector<float> adversarial(std::mt19937& rng, size_t n) {
  std::vector<float> v(n);
  std::uniform_int_distribution<uint32_t> any;
  const float specials[] = {0.f, -0.f, INFINITY, -INFINITY, NAN, 1e-45f, -1e-41f, 1.17549435e-38f, 3.4e38f, -3.4e38f,
                            240.f, 248.f, 255.9f, 256.f, 57344.f, 61440.f, 65504.f, 1.f, -1.f, 0.5f, 1.5f, 2.5f};
  for (size_t i = 0; i < n; ++i) v[i] = i < sizeof(specials) / 4 ? specials[i] : from_bits(any(rng));
  return v;
}

// compile-time casts (the ones the CUDA kernels inline) == the generic reference cast, on every upper-half bit
// pattern x a set of lower halves that exercise ties, just-above/below ties and all-ones mantissas
template <int E, int M>
static void check_static_casts() {
  const cpd::Format f = cpd::make_format(E, M, cpd::ROUND_REF);
  const uint32_t lows[] = {0x0000u, 0x0001u, 0x7FFFu, 0x8000u, 0x8001u, 0xFFFFu, 0x1234u, 0xFFF0u};
  long bad = 0;
  for (uint32_t hi = 0; hi < 65536u; ++hi)
    for (uint32_t lo : lows) {
      const float x = from_bits((hi << 16) | lo);
      const float want = cpd::quantize_ref(x, f);
      const float a = cpd::FormatT<E, M>::q(x), b = cpd::FormatT<E, M>::q_wide(x);
      if (!same(a, want) || !same(b, want)) {
        if (++bad <= 3) std::fprintf(stderr, "FAIL static cast (%d,%d) x=%a: q=%a q_wide=%a want=%a\n", E, M, x, a, b, want);
      }
    }
  // random mantissas over the exponents around and below the smallest normal (the inline subnormal path of q_wide)
  std::mt19937 rng(E * 100 + M);
  const int lo_e = cpd::FormatT<E, M>::lo;
  for (int i = 0; i < 2000000; ++i) {
    const uint32_t r = rng();
    int e = lo_e + 2 - int(r % 30u);
    if (e < 1) e = 1;
    if (e > 254) e = 254;
    const float x = from_bits((rng() & 0x807FFFFFu) | (uint32_t(e) << 23));
    const float want = cpd::quantize_ref(x, f);
    const float b = cpd::FormatT<E, M>::q_wide(x);
    if (!same(b, want) && ++bad <= 3) std::fprintf(stderr, "FAIL q_wide (%d,%d) x=%a: %a want %a\n", E, M, x, b, want);
  }
  if (bad) failures += (int)bad;
}

// the hardware pack path's algorithm: IEEE RNE convert of pretruncate(x) == reference cast, below the overflow band
template <int E, int M>
static void check_pretruncate(float limit) {
  const cpd::Format fr = cpd::make_format(E, M, cpd::ROUND_REF), fi = cpd::make_format(E, M, cpd::ROUND_IEEE);
  long bad = 0, n = 0;
  std::mt19937 rng(E * 1000 + M);
  auto one = [&](float x) {
    if (!(std::fabs(x) < limit)) return;                 // NaN / Inf / overflow band take the software path on the GPU too
    ++n;
    const float want = cpd::quantize_ref(x, fr);
    const float got = cpd::quantize_ieee(cpd::pretruncate<E, M>(x), fi);
    if (!same(got, want) && ++bad <= 5) std::fprintf(stderr, "FAIL pretruncate (%d,%d) x=%a: %a want %a\n", E, M, x, got, want);
  };
  const uint32_t lows[] = {0x0000u, 0x0001u, 0x7FFFu, 0x8000u, 0x8001u, 0xFFFFu, 0x1234u, 0xFFF0u};
  for (uint32_t hi = 0; hi < 65536u; ++hi)
    for (uint32_t lo : lows) one(from_bits((hi << 16) | lo));
  const int lo_e = cpd::FormatT<E, M>::lo;
  for (int i = 0; i < 4000000; ++i) {
    int e = lo_e + 3 - int(rng() % 32u);
    if (e < 0) e = 0;
    one(from_bits((rng() & 0x807FFFFFu) | (uint32_t(e) << 23)));
  }
  if (bad) failures += (int)bad;
  std::printf("pretruncate (%d,%d): %ld patterns below %g checked\n", E, M, n, limit);
}

int main() {
  check_pretruncate<4, 3>(248.0f);
  check_pretruncate<5, 2>(61440.0f);
  std::mt19937 rng(1234);
  check_static_casts<4, 3>(); check_static_casts<5, 2>(); check_static_casts<5, 10>(); check_static_casts<8, 7>();
  check_static_casts<8, 23>(); check_static_casts<3, 0>(); check_static_casts<2, 1>(); check_static_casts<8, 0>();
  check_static_casts<1, 1>(); check_static_casts<6, 9>(); check_static_casts<3, 4>(); check_static_casts<8, 22>();
  check_static_casts<7, 16>(); check_static_casts<5, 23>();
  const int fmts[][2] = {{4, 3}, {5, 2}, {3, 0}, {5, 10}, {8, 7}, {8, 23}, {2, 1}, {8, 0}, {1, 1}, {6, 9}, {3, 4}, {8, 22}};

  // ---- cast: idempotent, sign-symmetric, monotone on finite inputs, identity for (8,23); zero-size is legal ----
  for (auto& f : fmts) {
    for (int rounding = 0; rounding < 2; ++rounding) {
      const int e = f[0], m = f[1];
      std::vector<float> x = adversarial(rng, 4099), y(x.size()), z(x.size()), neg(x.size()), yn(x.size());
      cpd::quantize_cpu(x.data(), y.data(), (int64_t)x.size(), e, m, rounding);
      cpd::quantize_cpu(y.data(), z.data(), (int64_t)y.size(), e, m, rounding);
      for (size_t i = 0; i < x.size(); ++i) neg[i] = -x[i];
      cpd::quantize_cpu(neg.data(), yn.data(), (int64_t)neg.size(), e, m, rounding);
      for (size_t i = 0; i < x.size(); ++i) {
        // the one documented exception to idempotence: quirk A maps the top band to 2^(emax+1), whose cast is Inf
        const bool quirk_value = rounding == 0 && std::isfinite(y[i]) && std::isinf(z[i]);
        CHECK(same(y[i], z[i]) || quirk_value, "Q(Q(x)) != Q(x): fmt (%d,%d) r%d x=%a y=%a z=%a", e, m, rounding, x[i], y[i], z[i]);
        // (fp32 denormals cast to an unsigned zero in the reference mode, so zeros compare by value)
        CHECK(same(yn[i], -y[i]) || (yn[i] == 0.f && y[i] == 0.f), "Q(-x) != -Q(x): fmt (%d,%d) x=%a", e, m, x[i]);
        if (e == 8 && m == 23 && !std::isnan(x[i]) && std::fabs(x[i]) >= 1.17549435e-38f)
          CHECK(same(y[i], x[i]), "(8,23) must be the identity on normals: %a -> %a", x[i], y[i]);
      }
      // monotone on a sorted finite ramp
      std::vector<float> ramp(2048), qr(2048);
      for (int i = 0; i < 2048; ++i) ramp[i] = std::ldexp(1.0f + i / 2048.0f, (i % 40) - 20) * ((i / 40) % 2 ? 1.f : 1.f);
      std::sort(ramp.begin(), ramp.end());
      cpd::quantize_cpu(ramp.data(), qr.data(), 2048, e, m, rounding);
      for (int i = 1; i < 2048; ++i) CHECK(!(qr[i] < qr[i - 1]), "cast not monotone: fmt (%d,%d) %a -> %a after %a", e, m, ramp[i], qr[i], qr[i - 1]);
      cpd::quantize_cpu(x.data(), y.data(), 0, e, m, rounding);                   // n == 0
      for (int shift : {-126, -3, 0, 5, 126}) {                                   // fused 2^shift scale == scale then cast
        std::vector<float> a(257), b(257), c(257);
        for (auto& t : a) t = std::ldexp(std::uniform_real_distribution<float>(-2.f, 2.f)(rng), -shift / 2);
        cpd::quantize_scaled_cpu(a.data(), b.data(), 257, e, m, rounding, shift);
        for (size_t i = 0; i < a.size(); ++i) c[i] = a[i] * std::ldexp(1.0f, shift);
        cpd::quantize_cpu(c.data(), c.data(), 257, e, m, rounding);
        for (size_t i = 0; i < a.size(); ++i) CHECK(same(b[i], c[i]), "fused shift %d differs: %a vs %a", shift, b[i], c[i]);
      }
    }
  }

  // ---- exact GEMM: ragged shapes, transposed (strided) operands, (8,23) plain == sequential fp32 dot ----
  for (int M : {1, 7, 17}) for (int N : {1, 16, 19}) for (int K : {1, 8, 33}) {
    std::vector<float> A((size_t)M * K), Bt((size_t)N * K), C((size_t)M * N, -7.f), R((size_t)M * N);
    std::uniform_real_distribution<float> u(-1.f, 1.f);
    for (auto& t : A) t = u(rng);
    for (auto& t : Bt) t = u(rng);
    cpd::gemm_exact_cpu(A.data(), K, 1, Bt.data(), 1, K, C.data(), N, M, N, K, 8, 23, 0, false);     // B given as B^T (strides)
    for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) {
      float s = 0.f;
      for (int k = 0; k < K; ++k) { volatile float p = A[(size_t)i * K + k] * Bt[(size_t)j * K + k]; s = s + p; }
      R[(size_t)i * N + j] = s;
    }
    for (size_t i = 0; i < C.size(); ++i) CHECK(same(C[i], R[i]), "gemm_exact (8,23) != sequential dot at %zu: %a vs %a (M%d N%d K%d)", i, C[i], R[i], M, N, K);
    for (auto& f : {std::pair<int, int>{4, 3}, {5, 2}}) {
      cpd::gemm_exact_cpu(A.data(), K, 1, Bt.data(), 1, K, C.data(), N, M, N, K, f.first, f.second, 0, true);
      std::vector<float> q(C.size());
      cpd::quantize_cpu(C.data(), q.data(), (int64_t)C.size(), f.first, f.second, 0);
      for (size_t i = 0; i < C.size(); ++i) CHECK(same(C[i], q[i]) || std::isinf(q[i]), "low-precision GEMM result is not a member of the format: %a", C[i]);
      cpd::gemm_tile_cpu(A.data(), K, 1, Bt.data(), 1, K, C.data(), N, M, N, K, f.first, f.second, 1, false, 16, 8, 7);
      for (size_t i = 0; i < C.size(); ++i) CHECK(std::isfinite(C[i]), "tile GEMM produced a non-finite value from finite inputs");
    }
  }

  // ---- split-K of the tile model == the ordered quantised fold of the splits' own tile GEMMs ----
  {
    const int M = 5, N = 6, K = 100, ks = 16;                       // 7 slices -> 3 splits of 3, 3, 1 slices
    std::vector<float> A((size_t)M * K), Bt((size_t)N * K), C((size_t)M * N), P((size_t)M * N), S((size_t)M * N), T((size_t)M * N);
    std::uniform_real_distribution<float> u(-1.f, 1.f);
    for (auto& t : A) t = u(rng);
    for (auto& t : Bt) t = u(rng);
    for (auto& f : {std::pair<int, int>{5, 10}, {4, 3}, {8, 23}}) {
      cpd::gemm_tile_cpu(A.data(), K, 1, Bt.data(), 1, K, C.data(), N, M, N, K, f.first, f.second, 1, false, ks, 8, 7, 3);
      std::fill(S.begin(), S.end(), 0.f);
      const int k_per_split = 3 * ks;
      for (int k0 = 0; k0 < K; k0 += k_per_split) {
        const int kz = std::min(k_per_split, K - k0);
        cpd::gemm_tile_cpu(A.data() + k0, K, 1, Bt.data() + k0, 1, K, P.data(), N, M, N, kz, f.first, f.second, 1, false, ks, 8, 7, 1);
        for (size_t i = 0; i < S.size(); ++i) T[i] = S[i] + P[i];
        cpd::quantize_cpu(T.data(), S.data(), (int64_t)S.size(), f.first, f.second, 1);               // s = Q(s + part), IEEE
      }
      for (size_t i = 0; i < C.size(); ++i) CHECK(same(C[i], S[i]), "split-K tile GEMM != ordered fold of the splits at %zu: %a vs %a", i, C[i], S[i]);
      cpd::gemm_tile_cpu(A.data(), K, 1, Bt.data(), 1, K, P.data(), N, M, N, K, f.first, f.second, 1, false, ks, 8, 7, 64);   // more splits than slices
      cpd::gemm_tile_cpu(A.data(), K, 1, Bt.data(), 1, K, T.data(), N, M, N, K, f.first, f.second, 1, false, ks, 8, 7, 7);
      for (size_t i = 0; i < C.size(); ++i) CHECK(same(P[i], T[i]), "splits > slices must clamp to one slice per split at %zu", i);
    }
  }

  // ---- APS fold: W ranks, several tensors incl. all-zero / Inf / denormal ones; fp32 identity == plain sum ----
  {
    const int T = 5, W = 4;
    const int64_t seg[T + 1] = {0, 48, 64, 1088, 1104, 1408};
    std::vector<std::vector<float>> src(W, std::vector<float>(seg[T]));
    std::normal_distribution<float> g(0.f, 1.f);
    for (int r = 0; r < W; ++r)
      for (int64_t i = 0; i < seg[T]; ++i) src[r][i] = g(rng) * 1e-3f;
    for (int r = 0; r < W; ++r) {
      for (int64_t i = seg[1]; i < seg[2]; ++i) src[r][i] = 0.f;                 // all-zero tensor: shift guard
      src[r][seg[3] + 3] = r == 1 ? INFINITY : 1.f;                               // non-finite maximum: shift guard
      for (int64_t i = seg[4]; i < seg[5]; ++i) src[r][i] = g(rng) * 1e-41f;      // fp32 denormals: shift clamp
    }
    std::vector<const float*> ptr(W);
    for (int r = 0; r < W; ++r) ptr[r] = src[r].data();
    std::vector<float> out(seg[T]), out2(seg[T]);
    std::vector<int> shifts(T);
    for (auto& f : fmts) for (int kahan = 0; kahan < 2; ++kahan) for (int aps = 0; aps < 2; ++aps) {
      cpd::aps_fold_cpu(ptr.data(), W, out.data(), seg, T, f[0], f[1], 0, aps, kahan, (float)W, shifts.data());
      cpd::aps_fold_cpu(ptr.data(), W, out2.data(), seg, T, f[0], f[1], 0, aps, kahan, (float)W, nullptr);
      for (int64_t i = 0; i < seg[T]; ++i) CHECK(same(out[i], out2[i]), "fold is not deterministic at %lld", (long long)i);
      if (aps) { CHECK(shifts[1] == 0, "all-zero tensor must get shift 0, got %d", shifts[1]); CHECK(shifts[3] == 0, "Inf tensor must get shift 0, got %d", shifts[3]);
                 CHECK(shifts[4] <= 126 && shifts[4] >= -126, "shift clamp violated: %d", shifts[4]); }
      for (int64_t i = seg[1]; i < seg[2]; ++i) CHECK(out[i] == 0.f, "sum of zeros must be zero");
    }
    cpd::aps_fold_cpu(ptr.data(), W, out.data(), seg, T, 8, 23, 0, false, false, (float)W, nullptr);
    for (int64_t i = 0; i < seg[3]; ++i) {
      float s = 0.f;
      for (int r = 0; r < W; ++r) s += src[r][i];
      CHECK(same(out[i], s), "(8,23) fold != rank-ordered fp32 sum at %lld", (long long)i);
    }
  }

  // ---- LARS / SGD step: finite, master == model afterwards, zero gradient + zero wd leaves weights alone ----
  {
    const int T = 3;
    const int64_t seg[T + 1] = {0, 16, 400, 416};
    std::vector<float> w(seg[T]), model(seg[T]), gr(seg[T]), mom(seg[T], 0.f), wdm = {1.f, 1.f, 0.f};
    std::normal_distribution<float> g(0.f, 1.f);
    for (auto& t : w) t = g(rng);
    for (auto& t : gr) t = g(rng) * 0.01f;
    for (int use_lars = 0; use_lars < 2; ++use_lars) for (int nesterov = 0; nesterov < 2; ++nesterov) {
      std::vector<float> w0 = w;
      cpd::lars_step_cpu(w.data(), model.data(), gr.data(), mom.data(), seg, T, 0.1f, 0.9f, 1e-4f, 0.001f, 0.f, use_lars, nesterov, wdm.data());
      for (int64_t i = 0; i < seg[T]; ++i) { CHECK(std::isfinite(w[i]), "optimizer produced a non-finite weight"); CHECK(same(w[i], model[i]), "master != model copy"); }
      bool moved = false;
      for (int64_t i = 0; i < seg[T]; ++i) moved |= !same(w[i], w0[i]);
      CHECK(moved, "optimizer step did not change the weights");
    }
    std::vector<float> zero(seg[T], 0.f), m2(seg[T], 0.f), w1 = w;
    cpd::lars_step_cpu(w1.data(), nullptr, zero.data(), m2.data(), seg, T, 0.1f, 0.9f, 0.f, 0.001f, 0.f, false, false, nullptr);
    for (int64_t i = 0; i < seg[T]; ++i) CHECK(same(w1[i], w[i]), "zero gradient + zero weight decay must be a no-op");
  }

  std::printf("cpu_ops_selftest: %d failure(s)\n", failures);
  return failures ? 1 : 0;
}
