// wire_hw.cuh — hardware-assisted wire codec and fold cast for the 8-bit formats.
//
// The reference-exact cast has two quirks that IEEE converts do not reproduce:
// A (overflow tested before rounding) and B (truncating shift for target subnormals).  Inside the
// all-reduce fold, however, every value that is cast is an exact fp32 sum or difference of two
// MEMBERS of the format (multiples of its subnormal quantum), so a subnormal result needs no
// rounding at all and quirk B cannot trigger; quirk A only concerns |x| >= 2^(emax+1) - ulp/2.
// Hence for |x| below that bound the reference cast == the IEEE cast == Blackwell's
// F2FP.SATFINITE.{E4M3,E5M2} convert, and CPD-(4,3)/(5,2) codes below the all-ones exponent are
// bit-identical to OCP E4M3/E5M2 codes.  Everything else (NaN, Inf, the quirk band, reserved
// codes) takes the exact software path.  Result: bit-exact with the software implementation at
// ~1/3 of the instructions.
#pragma once
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include "fpformat.h"

namespace cpd {

// ---- generic (software) operations ----------------------------------------------------
template <class F>
struct FoldOps {
  static constexpr bool kHw = false;
  static __device__ __forceinline__ float qsum(const F& f, float x) { return f.q(x); }
  template <int N>
  static __device__ __forceinline__ void qsum_n(const F& f, float (&x)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = f.q(x[i]);
  }
  static __device__ __forceinline__ void decode4(const F& f, uint32_t w, float (&o)[4]) {
    o[0] = f.unpack(w & 0xFFu); o[1] = f.unpack((w >> 8) & 0xFFu); o[2] = f.unpack((w >> 16) & 0xFFu); o[3] = f.unpack(w >> 24);
  }
  static __device__ __forceinline__ uint32_t encode4(const F& f, const float (&x)[4]) {
    return f.pack(x[0]) | (f.pack(x[1]) << 8) | (f.pack(x[2]) << 16) | (f.pack(x[3]) << 24);
  }
  // N values: x <- Q(x * sc), reference cast (the first cast of a raw gradient: quirk B can trigger, so the fold's
  // exact-sum shortcut does not apply)
  template <int N>
  static __device__ __forceinline__ void qscaled_n(const F& f, float (&x)[N], float sc) {
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = f.q_wide(x[i] * sc);
  }
  // scale + cast + pack of one granule in one step: only the hardware formats have it
  static __device__ __forceinline__ bool pack16_scaled(const F&, const float (&)[16], float, uint32_t (&)[4]) { return false; }
  static __device__ __forceinline__ void decode16(const F& f, const uint32_t (&w)[4], float (&x)[16]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float o[4];
      decode4(f, w[k], o);
      x[4 * k] = o[0]; x[4 * k + 1] = o[1]; x[4 * k + 2] = o[2]; x[4 * k + 3] = o[3];
    }
  }
};

// ---- hardware paths -------------------------------------------------------------------------
template <int E_BITS, int M_BITS, __nv_fp8_interpretation_t INTERP>
struct FoldOpsFp8 {
  using F = FmtT<E_BITS, M_BITS>;
  static constexpr bool kHw = true;
  static constexpr int bias = (1 << (E_BITS - 1)) - 1;
  // first magnitude that must NOT take the hardware path: 2^(emax+1) - half ulp of the top binade
  static constexpr float kLimit = (E_BITS == 4) ? 248.0f : 61440.0f;
  static constexpr uint32_t kExpMask8 = ((1u << E_BITS) - 1u) << M_BITS;            // exponent bits of one code
  static constexpr uint32_t kExpMask32 = kExpMask8 * 0x01010101u;
  static constexpr uint32_t kCarry32 = (1u << M_BITS) * 0x01010101u;               // + this sets bit 7 iff exponent is all ones

  static __device__ __forceinline__ float qsum(const F& f, float x) {
    if (fabsf(x) < kLimit) {
      const __nv_fp8_storage_t c = __nv_cvt_float_to_fp8(x, __NV_SATFINITE, INTERP);
      return __half2float(__half(__nv_cvt_fp8_to_halfraw(c, INTERP)));
    }
    return f.q(x);                                    // NaN / Inf / quirk band: exact software cast
  }
  // out-of-line exact cast for the rare lanes of the wide version below (keeps its cold block to one CALL per element)
  static __device__ __noinline__ float qsum_cold(float x) { return FormatT<E_BITS, M_BITS>::q(x); }
  // N casts behind ONE guard: the NaN-propagating maximum of the magnitudes is below the limit iff every element is a
  // finite value below the quirk band, and then all N go through paired converts with no per-element control flow
  // (per element: 1 FMNMX + 1/2 F2FP pack + 1/2 F2FP unpack + 1 HADD2.F32, against ~11 instructions and one
  // data-dependent branch for the scalar form — the branch latency, not the ALU work, bounded a lone warp).
  template <int N>
  static __device__ __forceinline__ void qsum_n(const F& f, float (&x)[N]) {
    static_assert(N % 2 == 0, "pairs");
    float g = fabsf(x[0]);
#pragma unroll
    for (int i = 1; i < N; ++i) asm("max.NaN.f32 %0, %0, %1;" : "+f"(g) : "f"(fabsf(x[i])));
    if (g < kLimit) {                                 // false when any element is NaN
#pragma unroll
      for (int i = 0; i < N; i += 2) {
        const __nv_fp8x2_storage_t c = __nv_cvt_float2_to_fp8x2(make_float2(x[i], x[i + 1]), __NV_SATFINITE, INTERP);
        const float2 r = __half22float2(__half2(__nv_cvt_fp8x2_to_halfraw2(c, INTERP)));
        x[i] = r.x; x[i + 1] = r.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i) x[i] = qsum_cold(x[i]);
    }
  }
  // N values: x <- Q(x * sc) with the REFERENCE cast through the conversion hardware: pretruncate() clears the bits the
  // reference's truncating subnormal shift discards, after which the IEEE convert is the reference cast (proven on all
  // 2^32 patterns below the overflow band, tests/cpp); NaN / Inf / overflow-band groups take the software cast.
  template <int N>
  static __device__ __forceinline__ void qscaled_n(const F& f, float (&x)[N], float sc) {
    static_assert(N % 2 == 0, "pairs");
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] *= sc;
    float g = fabsf(x[0]);
#pragma unroll
    for (int i = 1; i < N; ++i) asm("max.NaN.f32 %0, %0, %1;" : "+f"(g) : "f"(fabsf(x[i])));
    if (g < kLimit) {
#pragma unroll
      for (int i = 0; i < N; i += 2) {
        const __nv_fp8x2_storage_t c = __nv_cvt_float2_to_fp8x2(
            make_float2(pretruncate<E_BITS, M_BITS>(x[i]), pretruncate<E_BITS, M_BITS>(x[i + 1])), __NV_SATFINITE, INTERP);
        const float2 r = __half22float2(__half2(__nv_cvt_fp8x2_to_halfraw2(c, INTERP)));
        x[i] = r.x; x[i + 1] = r.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i) x[i] = qsum_cold_wide(x[i]);
    }
  }
  static __device__ __noinline__ float qsum_cold_wide(float x) { return FormatT<E_BITS, M_BITS>::q_wide(x); }
  static __device__ __forceinline__ void decode4(const F& f, uint32_t w, float (&o)[4]) {
    if ((((w & kExpMask32) + kCarry32) & 0x80808080u) == 0u) {          // no reserved (all-ones exponent) code among the 4
      const __half2_raw lo = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(w & 0xFFFFu), INTERP);
      const __half2_raw hi = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(w >> 16), INTERP);
      const float2 a = __half22float2(__half2(lo)), b = __half22float2(__half2(hi));
      o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
    } else {
      o[0] = f.unpack(w & 0xFFu); o[1] = f.unpack((w >> 8) & 0xFFu); o[2] = f.unpack((w >> 16) & 0xFFu); o[3] = f.unpack(w >> 24);
    }
  }
  // Scale + reference cast + pack of one granule straight to wire codes: pretruncate() clears the bits the reference's
  // truncating subnormal shift discards, after which the RNE hardware convert IS the reference cast (fpformat.h).
  // ~10 instructions per element and no per-element control flow, against ~25 with branches for q() + encode4().
  // Returns false (nothing produced) if any scaled value is NaN or in the overflow band: the caller then takes the
  // exact software path for the granule.
  static __device__ __forceinline__ bool pack16_scaled(const F&, const float (&x)[16], float sc, uint32_t (&w)[4]) {
    float t[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) t[i] = x[i] * sc;
    float g = fabsf(t[0]);
#pragma unroll
    for (int i = 1; i < 16; ++i) asm("max.NaN.f32 %0, %0, %1;" : "+f"(g) : "f"(fabsf(t[i])));
    if (!(g < kLimit)) return false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(pretruncate<E_BITS, M_BITS>(t[4 * k]), pretruncate<E_BITS, M_BITS>(t[4 * k + 1])),
                                                   __NV_SATFINITE, INTERP);
      const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(pretruncate<E_BITS, M_BITS>(t[4 * k + 2]), pretruncate<E_BITS, M_BITS>(t[4 * k + 3])),
                                                   __NV_SATFINITE, INTERP);
      w[k] = lo | (hi << 16);
    }
    return true;
  }
  // one granule (16 codes) behind one reserved-code test
  static __device__ __forceinline__ void decode16(const F& f, const uint32_t (&w)[4], float (&x)[16]) {
    const uint32_t any = (((w[0] & kExpMask32) + kCarry32) | ((w[1] & kExpMask32) + kCarry32) |
                          ((w[2] & kExpMask32) + kCarry32) | ((w[3] & kExpMask32) + kCarry32)) & 0x80808080u;
    if (any == 0u) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 a = __half22float2(__half2(__nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(w[k] & 0xFFFFu), INTERP)));
        const float2 b = __half22float2(__half2(__nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(w[k] >> 16), INTERP)));
        x[4 * k] = a.x; x[4 * k + 1] = a.y; x[4 * k + 2] = b.x; x[4 * k + 3] = b.y;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int j = 0; j < 4; ++j) x[4 * k + j] = unpack_cold((w[k] >> (8 * j)) & 0xFFu);
      }
    }
  }
  static __device__ __noinline__ float unpack_cold(uint32_t c) { return wire_unpack(c, E_BITS, wire_man(M_BITS)); }
  static __device__ __forceinline__ uint32_t encode4(const F& f, const float (&x)[4]) {
    const float m = fmaxf(fmaxf(fabsf(x[0]), fabsf(x[1])), fmaxf(fabsf(x[2]), fabsf(x[3])));
    const bool nan = (x[0] != x[0]) | (x[1] != x[1]) | (x[2] != x[2]) | (x[3] != x[3]);
    if (m < kLimit && !nan) {                                             // all four are finite members below the reserved range
      const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(x[0], x[1]), __NV_SATFINITE, INTERP);
      const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(x[2], x[3]), __NV_SATFINITE, INTERP);
      return lo | (hi << 16);
    }
    return f.pack(x[0]) | (f.pack(x[1]) << 8) | (f.pack(x[2]) << 16) | (f.pack(x[3]) << 24);
  }
};

template <> struct FoldOps<FmtT<4, 3>> : FoldOpsFp8<4, 3, __NV_E4M3> {};
template <> struct FoldOps<FmtT<5, 2>> : FoldOpsFp8<5, 2, __NV_E5M2> {};

// fold steps whose casts act on exact sums / differences of format members (see header comment)
template <class F>
__device__ __forceinline__ void fold_plain_members(const F& f, float& s, float v) { s = FoldOps<F>::qsum(f, s + v); }
template <class F>
__device__ __forceinline__ void fold_kahan_members(const F& f, float& s, float& c, float v) {
  const float y = FoldOps<F>::qsum(f, v - c);
  const float t = FoldOps<F>::qsum(f, s + y);
  c = FoldOps<F>::qsum(f, FoldOps<F>::qsum(f, t - s) - y);
  s = t;
}

// the same two steps over N accumulators at once (one guard per cast stage instead of one per element)
template <class F, int N>
__device__ __forceinline__ void fold_plain_members_n(const F& f, float (&s)[N], const float (&v)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) s[i] += v[i];
  FoldOps<F>::template qsum_n<N>(f, s);
}
template <class F, int N>
__device__ __forceinline__ void fold_kahan_members_n(const F& f, float (&s)[N], float (&c)[N], const float (&v)[N]) {
  float y[N], t[N], d[N];
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = v[i] - c[i];
  FoldOps<F>::template qsum_n<N>(f, y);
#pragma unroll
  for (int i = 0; i < N; ++i) t[i] = s[i] + y[i];
  FoldOps<F>::template qsum_n<N>(f, t);
#pragma unroll
  for (int i = 0; i < N; ++i) d[i] = t[i] - s[i];
  FoldOps<F>::template qsum_n<N>(f, d);
#pragma unroll
  for (int i = 0; i < N; ++i) d[i] -= y[i];
  FoldOps<F>::template qsum_n<N>(f, d);
#pragma unroll
  for (int i = 0; i < N; ++i) { c[i] = d[i]; s[i] = t[i]; }
}

}  // namespace cpd
