// fpformat.h — (exp,man) custom floating-point format core, host + device.
//
// Numeric spec: SURVEY.md §2.9.1 (behaviour of the reference's cast_precision,
// /root/reference/CPDtorch/quant/quant_cuda/float_kernel.cu:10-92).  This is a
// fresh integer-on-the-bit-pattern formulation: one add + one mask for the
// target-normal case (the carry walks into the exponent field), a truncating
// shift + the same add/mask for target-subnormals, and an exact power-of-two
// rebuild.  Two rounding modes:
//   ROUND_REF  — bit-exact with the reference, including its quirks
//                (A: overflow test before rounding, B: truncating subnormal
//                 shift, C: man==23 never rounds subnormals)
//   ROUND_IEEE — textbook IEEE-754 RNE: sticky-aware subnormals, overflow test
//                after rounding.
// Everything here is constexpr-friendly and used by the CPU ops, the SIMT
// kernels, the tcgen05 GEMM epilogue and the fused all-reduce alike.
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>

#if defined(__CUDACC__)
#define CPD_HD __host__ __device__ __forceinline__
#define CPD_D __device__ __forceinline__
#define CPD_COLD __host__ __device__ __noinline__
#else
#define CPD_HD inline
#define CPD_D inline
#define CPD_COLD inline
#endif

namespace cpd {

enum Rounding : int { ROUND_REF = 0, ROUND_IEEE = 1 };

CPD_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u; std::memcpy(&u, &f, 4); return u;
#endif
}
CPD_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f; std::memcpy(&f, &u, 4); return f;
#endif
}

// Runtime description of a format; all derived constants precomputed once.
struct Format {
  int exp_bits;    // 1..8
  int man_bits;    // 0..23
  int bias;        // 2^(e-1) - 1
  int drop;        // 23 - man
  int ne_inf;      // 2^e - 1 : first biased exponent that is not finite
  uint32_t half_m1;  // (1 << (drop-1)) - 1   (0 when drop == 0)
  uint32_t keep;     // ~((1 << drop) - 1)
  uint32_t sub_scale_bits;  // fp32 bits of 2^(1 - bias - 23)   (0 when e == 8)
  int rounding;
};

CPD_HD Format make_format(int e, int m, int rounding = ROUND_REF) {
  Format f;
  f.exp_bits = e;
  f.man_bits = m;
  f.bias = (1 << (e - 1)) - 1;
  f.drop = 23 - m;
  f.ne_inf = (1 << e) - 1;
  f.half_m1 = f.drop > 0 ? ((1u << (f.drop - 1)) - 1u) : 0u;
  f.keep = f.drop > 0 ? ~((1u << f.drop) - 1u) : 0xFFFFFFFFu;
  int se = 1 - f.bias - 23 + 127;  // biased fp32 exponent of the subnormal unit
  f.sub_scale_bits = se > 0 ? (uint32_t(se) << 23) : 0u;
  f.rounding = rounding;
  return f;
}

CPD_HD bool is_identity(int e, int m) { return e == 8 && m == 23; }

// RNE of an integer significand at bit position `drop` (add-and-mask).
CPD_HD uint32_t rne_bits(uint32_t a, int drop, uint32_t half_m1, uint32_t keep) {
  return (a + half_m1 + ((a >> drop) & 1u)) & keep;
}
// Same on a full fp32 bit pattern (exponent field included).  With man == 0 the kept
// LSB is the implicit leading one — always odd — so ties round up; bit 23 of the
// pattern is the exponent's LSB and must not be used.
CPD_HD uint32_t rne_pattern(uint32_t a, int drop, uint32_t half_m1, uint32_t keep) {
  const uint32_t lsb = drop == 23 ? 1u : ((a >> drop) & 1u);
  return (a + half_m1 + lsb) & keep;
}

// ---- reference-exact cast --------------------------------------------------
CPD_HD float quantize_ref(float x, const Format& f) {
  const uint32_t u = f2u(x);
  const uint32_t a = u & 0x7FFFFFFFu;
  const uint32_t sign = u & 0x80000000u;
  const uint32_t E = a >> 23;
  if (E == 255u || a == 0u) return x;          // Inf / NaN / +-0 pass through
  if (E == 0u) return 0.0f;                     // fp32 denormal -> +0
  const int ne = int(E) - 127 + f.bias;
  if (ne >= f.ne_inf) return u2f(sign | 0x7F800000u);   // overflow BEFORE rounding (quirk A)
  if (ne >= 1) {
    if (f.drop == 0) return x;
    return u2f(sign | rne_pattern(a, f.drop, f.half_m1, f.keep));
  }
  // target-subnormal: truncating shift (quirk B), then RNE at the same position
  int s = 1 - ne;
  if (s > 31) s = 31;
  uint32_t M = ((a & 0x7FFFFFu) | 0x800000u) >> s;
  if (f.drop > 0) M = rne_bits(M, f.drop, f.half_m1, f.keep);   // quirk C: m==23 keeps M
  const float r = float(M) * u2f(f.sub_scale_bits);      // exact: M < 2^24, power-of-two scale
  return u2f(f2u(r) | sign);
}

// ---- IEEE-style cast (RNE with sticky, overflow after rounding) ------------
CPD_HD float quantize_ieee(float x, const Format& f) {
  const uint32_t u = f2u(x);
  const uint32_t a = u & 0x7FFFFFFFu;
  const uint32_t sign = u & 0x80000000u;
  const uint32_t E = a >> 23;
  if (E == 255u || a == 0u) return x;
  const int emin = 1 - f.bias;                       // smallest normal exponent
  const int te = (E == 0u) ? -126 : int(E) - 127;
  uint32_t r;
  if (E != 0u && te >= emin) {
    r = (f.drop == 0) ? a : rne_pattern(a, f.drop, f.half_m1, f.keep);
  } else {
    // subnormal in the target: quantum is 2^(emin - man).  Scale so the quantum
    // is 1.0, round with the FPU's RNE, scale back.  Both scalings are exact.
    if (f.exp_bits == 8) {
      // fp32 denormal input and e == 8: quantum 2^(-126-m) <= fp32's own; keep bits
      if (f.drop == 0) return x;
      r = rne_pattern(a, f.drop, f.half_m1, f.keep);
    } else {
      const int q = emin - f.man_bits;               // exponent of the quantum
      const float up = u2f(uint32_t(127 - q) << 23); // 2^-q  (q >= -63-23, so 127-q <= 213)
      const float dn = u2f(uint32_t(127 + q) << 23); // 2^q
#if defined(__CUDA_ARCH__)
      const float v = rintf(u2f(a) * up) * dn;
#else
      const float v = std::nearbyintf(u2f(a) * up) * dn;
#endif
      r = f2u(v);
    }
  }
  // overflow AFTER rounding: anything at or above 2^(emax+1) becomes Inf
  const uint32_t inf_thresh = uint32_t(f.ne_inf - f.bias + 127);   // biased fp32 exponent of 2^(emax+1)
  if (inf_thresh < 255u && (r >> 23) >= inf_thresh) r = 0x7F800000u;
  return u2f(sign | r);
}

CPD_HD float quantize(float x, const Format& f) {
  return f.rounding == ROUND_REF ? quantize_ref(x, f) : quantize_ieee(x, f);
}

// ---- compile-time specialised reference-exact cast (hot loops) -------------
// Fast path first: target-normal inputs take one compare on the exponent
// field, one add, one and.  Everything else goes through the general routine.
template <int E_BITS, int M_BITS>
struct FormatT {
  static constexpr int bias = (1 << (E_BITS - 1)) - 1;
  static constexpr int drop = 23 - M_BITS;
  static constexpr int ne_inf = (1 << E_BITS) - 1;
  static constexpr uint32_t half_m1 = drop > 0 ? ((1u << (drop > 0 ? drop - 1 : 0)) - 1u) : 0u;
  static constexpr uint32_t keep = drop > 0 ? ~((1u << drop) - 1u) : 0xFFFFFFFFu;
  // fp32 biased exponent range that is target-normal: [lo, hi)
  static constexpr int lo = 1 - bias + 127 > 1 ? 1 - bias + 127 : 1;
  static constexpr int hi = ne_inf - bias + 127 < 255 ? ne_inf - bias + 127 : 255;
  static constexpr int sub_se = 1 - bias - 23 + 127;

  // cold path (specials, overflow, target-subnormals): kept out of line so the hot
  // loops that inline q() stay small
  static CPD_COLD float q_slow(float x) {
    const uint32_t u = f2u(x);
    const uint32_t a = u & 0x7FFFFFFFu;
    const uint32_t E = a >> 23;
    if (E == 255u || a == 0u) return x;
    if (E == 0u) return 0.0f;
    if (int(E) >= hi) return u2f((u & 0x80000000u) | 0x7F800000u);
    int s = lo - int(E);
    if (s > 31) s = 31;
    uint32_t M = ((a & 0x7FFFFFu) | 0x800000u) >> s;
    if (drop > 0) M = (M + half_m1 + ((M >> drop) & 1u)) & keep;
    const float r = float(M) * u2f(sub_se > 0 ? (uint32_t(sub_se) << 23) : 0u);
    return u2f(f2u(r) | (u & 0x80000000u));
  }
  static CPD_HD float q(float x) {
    const uint32_t u = f2u(x);
    const uint32_t a = u & 0x7FFFFFFFu;
    const uint32_t E = a >> 23;
    if (E - uint32_t(lo) < uint32_t(hi - lo)) {            // target-normal fast path
      if (drop == 0) return x;
      const uint32_t lsb = drop == 23 ? 1u : ((a >> drop) & 1u);   // man == 0: implicit one is the LSB
      const uint32_t r = (a + half_m1 + lsb) & keep;
      return u2f((u & 0x80000000u) | r);
    }
    return q_slow(x);
  }
  // Same cast with the target-SUBNORMAL range (and zeros) handled inline as well.  APS maps the largest |g| of
  // a tensor to the top of the format, so with real gradients (several decades of dynamic range) a large share
  // of the scaled values land below the smallest normal of a 4- or 5-bit-exponent format: in q() every warp
  // then takes the out-of-line call for almost every element.  Here they cost ~12 straight-line integer
  // instructions (truncating shift = quirk B, RNE at the same bit position, exact rebuild through the 2^23
  // magic constant: no int->float conversion).  Only Inf/NaN/overflow still leave the inlined code.
  static CPD_HD float q_wide(float x) {
    const uint32_t u = f2u(x);
    const uint32_t a = u & 0x7FFFFFFFu;
    const uint32_t E = a >> 23;
    if (E - uint32_t(lo) < uint32_t(hi - lo)) {            // target-normal
      if (drop == 0) return x;
      const uint32_t lsb = drop == 23 ? 1u : ((a >> drop) & 1u);
      const uint32_t r = (a + half_m1 + lsb) & keep;
      return u2f((u & 0x80000000u) | r);
    }
    if (E < uint32_t(lo)) {
      if (E == 0u) return u2f(a == 0u ? u : 0u);           // +-0 passes through, fp32 denormals become +0
      uint32_t s = uint32_t(lo) - E;                        // >= 1
      if (s > 31u) s = 31u;
      uint32_t M = ((a & 0x7FFFFFu) | 0x800000u) >> s;      // < 2^23
      if (drop > 0) M = (M + half_m1 + ((M >> drop) & 1u)) & keep;     // <= 2^23
      const float r = (u2f(0x4B000000u + M) - 8388608.0f) * u2f(sub_se > 0 ? (uint32_t(sub_se) << 23) : 0u);
      return u2f(f2u(r) | (u & 0x80000000u));
    }
    return q_slow(x);
  }
};

// ---- accumulation steps ----------------------------------------------------
// Plain:  s = Q(s + v)
// Kahan:  y = Q(v - c); t = Q(s + y); c = Q(Q(t - s) - y); s = t
// (reference: dist_util.py:60-89 for the gradient fold, float_kernel.cu:181-195
//  for the GEMM inner step)
template <class Q>
CPD_HD void fold_plain(const Q& q, float& s, float v) { s = q.q(s + v); }

template <class Q>
CPD_HD void fold_kahan(const Q& q, float& s, float& c, float v) {
  const float y = q.q(v - c);
  const float t = q.q(s + y);
  c = q.q(q.q(t - s) - y);
  s = t;
}

// ---- APS (auto precision scaling) -----------------------------------------
// Reference: dist_util.py:26-34.  me = ceil(log2(max|g * W|)) evaluated in fp32
// exactly like torch does, MAX over ranks, shift = bias - me.  Guards that the
// reference lacks: all-zero tensors and non-finite maxima give shift 0; the
// shift is clamped so that 2^shift and 2^-shift are normal fp32 numbers.
CPD_HD float aps_local_exponent(float amax_abs, float world) {
  const float v = amax_abs * world;
#if defined(__CUDA_ARCH__)
  return ceilf(log2f(v));
#else
  return std::ceil(std::log2(v));
#endif
}
CPD_HD int aps_shift_from_exponent(float me, int exp_bits) {
  const int bias = (1 << (exp_bits - 1)) - 1;
  if (!(me > -1e30f) || !(me < 1e30f)) return 0;     // -inf (all zero), +inf, NaN
  int s = bias - int(me);
  if (s > 126) s = 126;
  if (s < -126) s = -126;
  return s;
}
CPD_HD float pow2i(int s) { return u2f(uint32_t(127 + s) << 23); }   // s in [-126,127]

// ---- wire format -----------------------------------------------------------
// Already-quantized values are packed to sign|biased exponent|mantissa in
// 8 or 16 bits (fp32 otherwise).  The all-ones exponent is reserved:
// mantissa 0 = Inf, 1 = the finite quirk-A value 2^(emax+1), >=2 = NaN.
// Formats whose code space has no room for the three specials are bumped to the
// next width.
// The codec always carries at least 2 mantissa bits so the three specials fit.
CPD_HD constexpr int wire_man(int m) { return m < 2 ? 2 : m; }
CPD_HD int wire_bits(int e, int m) {
  const int need = 1 + e + wire_man(m);
  if (e < 2) return 32;
  if (need <= 8) return 8;
  if (need <= 16) return 16;
  return 32;
}

static CPD_COLD uint32_t wire_pack_special(uint32_t a, uint32_t top, uint32_t ne_inf, int M_BITS) {
  uint32_t code = ne_inf << M_BITS;
  if (a == 0x7F800000u) {}                  // Inf
  else if (a == top) code |= 1u;            // finite 2^(emax+1) (quirk A)
  else code |= 2u;                          // NaN (or a non-member: poisoned)
  return code;
}
static CPD_COLD float wire_unpack_special(uint32_t s, uint32_t mant, uint32_t top_bits) {
  if (mant == 0u) return u2f(s | 0x7F800000u);
  if (mant == 1u) return u2f(s | top_bits);
  return u2f(s | 0x7FC00000u);
}

CPD_HD uint32_t wire_pack(float v, const int E_BITS, const int M_BITS) {
  const int bias = (1 << (E_BITS - 1)) - 1;
  const int ne_inf = (1 << E_BITS) - 1;
  const uint32_t sign_pos = E_BITS + M_BITS;
  const uint32_t u = f2u(v);
  const uint32_t a = u & 0x7FFFFFFFu;
  const uint32_t s = (u >> 31) << sign_pos;
  const uint32_t top = uint32_t(ne_inf - bias + 127) << 23;   // bits of 2^(emax+1)
  if (a >= top) return s | wire_pack_special(a, top, uint32_t(ne_inf), M_BITS);
  // scale by 2^(bias-127): the fp32 exponent field becomes the target's biased
  // exponent, target subnormals land in fp32 denormals (exact; needs -ftz=false).
  const float scaled = u2f(a) * u2f(uint32_t(bias) << 23);
  return s | (f2u(scaled) >> (23 - M_BITS));
}

CPD_HD float wire_unpack(uint32_t code, const int E_BITS, const int M_BITS) {
  const int bias = (1 << (E_BITS - 1)) - 1;
  const uint32_t ne_inf = (1u << E_BITS) - 1u;
  const uint32_t sign_pos = E_BITS + M_BITS;
  const uint32_t s = ((code >> sign_pos) & 1u) << 31;
  const uint32_t body = code & ((1u << sign_pos) - 1u);
  if ((body >> M_BITS) == ne_inf)
    return wire_unpack_special(s, body & ((1u << M_BITS) - 1u), uint32_t(int(ne_inf) - bias + 127) << 23);
  const float scaled = u2f(body << (23 - M_BITS)) * u2f(uint32_t(254 - bias) << 23);  // * 2^(127-bias)
  return u2f(f2u(scaled) | s);
}

// x with the low significand bits cleared that the reference's truncating subnormal shift discards (quirk B) and
// fp32 denormals turned into +0: an IEEE round-to-nearest-even convert of the result — e.g. Blackwell's
// F2FP.SATFINITE.E4M3/E5M2 — then reproduces the reference cast exactly for every finite |x| below the overflow
// band (the reference rounds the truncated significand at the same bit position; after the masking the hardware sees
// the same bits).  6 integer instructions, no control flow.  Proven against quantize_ref() in tests/cpp.
template <int E_BITS, int M_BITS>
CPD_HD float pretruncate(float x) {
  constexpr uint32_t lo = uint32_t(FormatT<E_BITS, M_BITS>::lo);
  const uint32_t u = f2u(x);
  const uint32_t E = (u >> 23) & 0xFFu;
  uint32_t s = lo > E ? lo - E : 0u;                    // bits the reference shifts out (0 for target-normal inputs)
  if (s > 23u) s = 23u;                                 // whole mantissa gone: +-2^(E-127) rounds to +-0 like the reference's M = 0
  const uint32_t t = u & (0xFFFFFFFFu << s);
  return u2f((E == 0u && (u & 0x7FFFFFu) != 0u) ? 0u : t);
}

// Format policies used by templated kernels: q() = cast, pack()/unpack() = wire codec.
template <int E_BITS, int M_BITS>
struct FmtT {
  static constexpr int kExp = E_BITS, kMan = M_BITS;
  CPD_HD float q(float x) const { return FormatT<E_BITS, M_BITS>::q(x); }
  CPD_HD float q_wide(float x) const { return FormatT<E_BITS, M_BITS>::q_wide(x); }   // subnormal range inline too
  CPD_HD uint32_t pack(float v) const { return wire_pack(v, E_BITS, wire_man(M_BITS)); }
  CPD_HD float unpack(uint32_t c) const { return wire_unpack(c, E_BITS, wire_man(M_BITS)); }
  CPD_HD int exp_bits() const { return E_BITS; }
};
// Runtime (exp, man): the generic, slower policy.  Its members stay out of line on the
// device so kernels that are unrolled over 16 elements x W ranks do not explode.
static CPD_COLD float quantize_rt(float x, const Format& f) { return quantize(x, f); }
static CPD_COLD uint32_t wire_pack_rt(float v, int e, int m) { return wire_pack(v, e, m); }
static CPD_COLD float wire_unpack_rt(uint32_t c, int e, int m) { return wire_unpack(c, e, m); }
struct FmtR {
  Format f;
  CPD_HD float q(float x) const { return quantize_rt(x, f); }
  CPD_HD float q_wide(float x) const { return quantize_rt(x, f); }
  CPD_HD uint32_t pack(float v) const { return wire_pack_rt(v, f.exp_bits, wire_man(f.man_bits)); }
  CPD_HD float unpack(uint32_t c) const { return wire_unpack_rt(c, f.exp_bits, wire_man(f.man_bits)); }
  CPD_HD int exp_bits() const { return f.exp_bits; }
};

}  // namespace cpd
