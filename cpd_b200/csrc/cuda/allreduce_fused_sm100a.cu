// allreduce_fused_sm100a.cu — the data-parallel gradient all-reduce of CPD as ONE
// sm_100a kernel per bucket, over NVSwitch peer memory.
//
// Reference path being replaced (CPDtorch/utils/dist_util.py:22-89): per parameter
// {abs, mul, max, log2, ceil} + host syncs + NCCL all_reduce(MAX) + scale +
// float_quantize + NCCL all_gather of fp32 + W x {add, memset, quantize} (or
// 4 quantizes + 4 add/sub for Kahan) + unscale: ~46 launches, 2 collectives and
// >= 2 host syncs per parameter, fp32 on the wire ((W-1)*4 B per element in).
//
// This kernel fuses all of it for a whole bucket (many tensors, per-tensor shift):
//   A  per-tensor max|g|                      -> grid barrier -> exponents PUSHED into every rank's mailbox,
//      signal peers; wait peers; shift[t] = bias - max_r ceil(log2(amax_r * W)) from the local mailbox
//   B  g -> Q(g * 2^shift) -> packed wire code (8/16/32 bit) into own symmetric
//      buffer                                  -> grid barrier -> signal peers; wait peers
//   C  one-shot : every rank loads all W peers' codes (128-bit peer loads over
//                 NVLink), folds in rank order with a cast after every add (Kahan:
//                 (res, c) in registers), unscales, stores fp32 gradients.
//      two-shot : rank r folds only the chunks it owns, stores the fp32 result
//                 locally and pushes the packed reduced codes into every peer's
//                 gather buffer (P2P stores, or ONE multimem.st through the NVLS
//                 multicast mapping) -> grid barrier -> signal; wait -> every rank
//                 unpacks + unscales the chunks it does not own.
// Every rank folds in the same rank order, so results are bit-identical on all
// ranks and equal to the reference semantics.  Buffers and
// mailboxes are double-buffered by epoch parity and flags are monotonic epochs, so
// no trailing barrier and no host-side reset is needed (CUDA-graph replay safe).
#include "../api.h"
#include "../common/fpformat.h"
#include "../common/device_utils.cuh"
#include "../common/wire_hw.cuh"
#include "format_dispatch.cuh"

namespace cpd {

constexpr int kArThreads = 256;
constexpr int kArMaxW = 8;
constexpr int kArMaxT = 1024;       // tensors per bucket (shifts live in shared memory)
constexpr int kFlagStride = 16;     // flags[phase][src]

struct ArParams {
  int rank, world;
  unsigned char* data[kArMaxW];
  unsigned char* gather[kArMaxW];
  float* mailbox[kArMaxW];
  unsigned* flags[kArMaxW];
  unsigned char* mc_gather;
  unsigned char* mc_data;
  long long cap_bytes;
  int max_tensors;
  float* grad;
  const Chunk* chunks;
  int n_chunks, T;
  int use_aps, algo, prequant, ident;
  int pack_hw;           // opt-in (algo bit 8): hardware scale+cast+pack of phase B for the fp8-shaped formats
  int poll_cta0;         // opt-in (algo bit 9): only CTA 0 polls the peers' flags, the others poll a local go-flag
  int weak_loads;        // opt-in (algo bit 10): plain instead of system-scope strong loads for the peers' wire data
  int no_extra_fence;    // opt-in (algo bit 11): no __threadfence_system() in front of the release store of a flag
  unsigned* ws;          // [0],[1] grid barrier; [2] epoch; [3] status; [4..4+2*max_tensors) amax[2][max_tensors]
  int* shifts_out;
  unsigned long long timeout;
};

template <class F> struct IsE5M2 { static constexpr bool value = false; };
template <> struct IsE5M2<FmtT<5, 2>> { static constexpr bool value = true; };

// ---- wire granule (16 elements) load / store / codec ------------------------------
template <int WIRE> struct Granule;
template <> struct Granule<8> { uint4 v[1]; };
template <> struct Granule<16> { uint4 v[2]; };
template <> struct Granule<32> { uint4 v[4]; };

template <int WIRE>
__device__ __forceinline__ Granule<WIRE> granule_load_sys(const unsigned char* base, long long elem) {
  Granule<WIRE> g;
  const unsigned char* p = base + elem * (WIRE / 8);
#pragma unroll
  for (int i = 0; i < WIRE / 8; ++i) g.v[i] = ld_sys_u4(p + 16 * i);
  return g;
}
// same through plain (weak) loads: legal after the acquire of the data flag, which orders them; the L1 cannot hold a
// stale copy because every wire buffer is read once per launch and the L1 is invalidated at launch boundaries
template <int WIRE>
__device__ __forceinline__ Granule<WIRE> granule_load_weak(const unsigned char* base, long long elem) {
  Granule<WIRE> g;
  const uint4* p = reinterpret_cast<const uint4*>(base + elem * (WIRE / 8));
#pragma unroll
  for (int i = 0; i < WIRE / 8; ++i) g.v[i] = p[i];
  return g;
}
template <int WIRE>
__device__ __forceinline__ void granule_store(unsigned char* base, long long elem, const Granule<WIRE>& g, bool sys) {
  unsigned char* p = base + elem * (WIRE / 8);
#pragma unroll
  for (int i = 0; i < WIRE / 8; ++i) {
    if (sys) st_sys_u4(p + 16 * i, g.v[i]);
    else st_u4(p + 16 * i, g.v[i]);
  }
}
template <int WIRE>
__device__ __forceinline__ void granule_store_mc(unsigned char* mc_base, long long elem, const Granule<WIRE>& g) {
  unsigned char* p = mc_base + elem * (WIRE / 8);
#pragma unroll
  for (int i = 0; i < WIRE / 8; ++i) multimem_st_u4(p + 16 * i, g.v[i]);
}

template <int WIRE, class F>
__device__ __forceinline__ void granule_decode(const F& fmt, const Granule<WIRE>& g, float (&x)[16]) {
  if (WIRE == 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      x[4 * i + 0] = __uint_as_float(g.v[i].x); x[4 * i + 1] = __uint_as_float(g.v[i].y);
      x[4 * i + 2] = __uint_as_float(g.v[i].z); x[4 * i + 3] = __uint_as_float(g.v[i].w);
    }
  } else if (WIRE == 16) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const unsigned w[4] = {g.v[i].x, g.v[i].y, g.v[i].z, g.v[i].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        x[8 * i + 2 * k] = fmt.unpack(w[k] & 0xFFFFu);
        x[8 * i + 2 * k + 1] = fmt.unpack(w[k] >> 16);
      }
    }
  } else {
    const unsigned w[4] = {g.v[0].x, g.v[0].y, g.v[0].z, g.v[0].w};
    FoldOps<F>::decode16(fmt, w, x);              // hardware E4M3/E5M2 unpack where the codes allow it
  }
}
template <int WIRE, class F>
__device__ __forceinline__ Granule<WIRE> granule_encode(const F& fmt, const float (&x)[16]) {
  Granule<WIRE> g;
  if (WIRE == 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      g.v[i] = make_uint4(__float_as_uint(x[4 * i]), __float_as_uint(x[4 * i + 1]), __float_as_uint(x[4 * i + 2]),
                          __float_as_uint(x[4 * i + 3]));
  } else if (WIRE == 16) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      unsigned w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) w[k] = fmt.pack(x[8 * i + 2 * k]) | (fmt.pack(x[8 * i + 2 * k + 1]) << 16);
      g.v[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
  } else {
    unsigned w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float in4[4] = {x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]};
      w[k] = FoldOps<F>::encode4(fmt, in4);
    }
    g.v[0] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  return g;
}

__device__ __forceinline__ void load16(const float* p, float (&x)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 v = *reinterpret_cast<const float4*>(p + 4 * i);
    x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
  }
}
__device__ __forceinline__ void store16(float* p, const float (&x)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(p + 4 * i) = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
}

// ---- cross-GPU signalling -----------------------------------------------------
// CTA 0 publishes `epoch` for `phase` into every peer's signal pad.
__device__ __forceinline__ void signal_peers(const ArParams& p, int phase, unsigned epoch) {
  if (blockIdx.x == 0 && threadIdx.x < p.world) {
    // st.release.sys already orders every earlier write (cumulatively) before the flag; the explicit fence in front of
    // it is belt and braces and can be dropped with the opt-in bit (A/B for the flag-round latency)
    if (!p.no_extra_fence) __threadfence_system();
    st_release_sys(p.flags[threadIdx.x] + phase * kFlagStride + p.rank, epoch);
  }
}
// every CTA waits until all peers have published `epoch` for `phase` (bounded spin).
// Default: W threads of EVERY CTA poll the peers' flags.  Opt-in (poll_cta0, algo bit 9): only CTA 0 polls the line
// the peers' release-stores land in and republishes a local go-flag (release/acquire at gpu scope) that the other
// CTAs poll - 296 x W pollers no longer compete with the incoming NVLink writes for that one L2 line.
// Returns true (CTA-uniform) when the wait timed out - now or in an earlier launch whose status the host has not
// cleared yet: the caller then POISONS its outputs (NaN) instead of folding stale peer data into the gradients, so a
// straggler can never silently corrupt a step (NCCL would have blocked; we fail loudly and bounded).
__device__ __forceinline__ bool wait_peers(const ArParams& p, int phase, unsigned epoch) {
  __shared__ unsigned s_bad;
  unsigned* go = p.ws + 4 + 2 * p.max_tensors + 26 + phase;
  if (!p.poll_cta0 || blockIdx.x == 0) {
    if (threadIdx.x < p.world) {
      const unsigned* f = p.flags[p.rank] + phase * kFlagStride + threadIdx.x;
      const long long t0 = clock64();
      while ((int)(ld_acquire_sys(f) - epoch) < 0) {
        if ((unsigned long long)(clock64() - t0) > p.timeout) {
          atomicExch(&p.ws[3], 0x80000000u | (unsigned(phase) << 8) | threadIdx.x);   // who is missing
          break;
        }
      }
    }
    if (p.poll_cta0) {
      __syncthreads();
      if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(go), "r"(epoch) : "memory");
    }
  } else if (threadIdx.x == 0) {
    const long long t0 = clock64();
    while ((int)(ld_acquire_gpu(go) - epoch) < 0) {
      if ((unsigned long long)(clock64() - t0) > 2 * p.timeout) {              // CTA 0 reports the missing peer
        atomicOr(&p.ws[3], 0x80000000u);
        break;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) s_bad = ld_acquire_gpu(&p.ws[3]) & 0x80000000u;
  __syncthreads();
  return s_bad != 0u;
}

// phase timestamps (ns, %globaltimer) of CTA 0 for attribution: ws tail, 8 x u64 after the amax buffers
__device__ __forceinline__ void stamp(const ArParams& p, int i) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    reinterpret_cast<unsigned long long*>(p.ws + 4 + 2 * p.max_tensors)[i] = t;
  }
}

// phase B for one granule: x -> Q(x * sc) -> wire codes in the rank's own buffer
template <int WIRE, class F>
__device__ __forceinline__ void pack_granule(const ArParams& p, const F& fmt, float (&x)[16], float sc, unsigned char* my_data, long long e0) {
  if (WIRE == 8 && p.pack_hw && (p.use_aps || p.prequant)) {
    unsigned w[4];
    if (FoldOps<F>::pack16_scaled(fmt, x, sc, w)) {
      st_u4(my_data + e0, make_uint4(w[0], w[1], w[2], w[3]));
      return;
    }
  }
  if (p.use_aps || p.prequant) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fmt.q_wide(x[i] * sc);
  }
  granule_store<WIRE>(my_data, e0, granule_encode<WIRE>(fmt, x), false);
}

template <class F, int WIRE, bool KAHAN>
__global__ void __launch_bounds__(kArThreads, 2)
allreduce_fused_kernel(const ArParams p, const F fmt) {
  __shared__ unsigned red[32];
  __shared__ int s_shift[kArMaxT];
  const int W = p.world, rank = p.rank;
  const unsigned epoch = p.ws[2] + 1u;
  const unsigned par = epoch & 1u;
  unsigned* bar = p.ws;
  unsigned* amax = p.ws + 4 + par * p.max_tensors;
  unsigned char* my_data = p.data[rank] + par * p.cap_bytes;

  bool poison = false;                 // set when a peer never showed up: outputs become NaN (see wait_peers)
  const float kPoison = __int_as_float(0x7FC00000);
  stamp(p, 0);
  // ---------------- single-round variant for small buckets (ALGO_ONESHOT_RAW) ---------------------------
  // Latency-bound messages: ONE flag round instead of two.  Each rank publishes its RAW fp32 gradients plus
  // its per-tensor exponents; after the single barrier every rank knows the global shift and applies
  // Q(g_r * 2^shift) to each peer's values on the fly while folding in rank order — the same casts in the
  // same order as the packed variants, hence bit-identical, at 4 B/element on the wire (fine below ~256 KB).
  if (p.algo == ALGO_ONESHOT_RAW) {
    float* my_raw = reinterpret_cast<float*>(my_data);
    for (int c = blockIdx.x; c < p.n_chunks; c += gridDim.x) {
      const Chunk ck = p.chunks[c];
      unsigned m = 0u;
      for (int g = threadIdx.x; g < (ck.len >> 4); g += kArThreads) {
        const uint4* q = reinterpret_cast<const uint4*>(p.grad + ck.start + 16 * g);
        uint4* d = reinterpret_cast<uint4*>(my_raw + ck.start + 16 * g);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 x = q[i];
          d[i] = x;
          m = max(max(m, x.x & 0x7FFFFFFFu), max(x.y & 0x7FFFFFFFu, max(x.z & 0x7FFFFFFFu, x.w & 0x7FFFFFFFu)));
        }
      }
      m = block_max_u32(m, red);
      if (threadIdx.x == 0 && m) atomicMax(&amax[ck.tensor], m);
    }
    grid_barrier(bar);
    stamp(p, 1);
    if (blockIdx.x == 0) {
      unsigned* amax_next = p.ws + 4 + (par ^ 1u) * p.max_tensors;
      for (int t = threadIdx.x; t < p.T; t += kArThreads) {
        const unsigned ab = amax[t];
        const float me = ab >= 0x7F800000u ? INFINITY : aps_local_exponent(__uint_as_float(ab), (float)W);
        for (int r = 0; r < W; ++r)
          asm volatile("st.relaxed.sys.global.f32 [%0], %1;" :: "l"(p.mailbox[r] + ((size_t)par * kArMaxW + rank) * p.max_tensors + t), "f"(me) : "memory");
      }
      for (int t = threadIdx.x; t < p.max_tensors; t += kArThreads) amax_next[t] = 0u;
      __syncthreads();
    }
    signal_peers(p, 1, epoch);
    poison |= wait_peers(p, 1, epoch);
    stamp(p, 4);
    for (int t = threadIdx.x; t < p.T; t += kArThreads) {
      float me = -INFINITY;
      for (int r = 0; r < W; ++r) {
        const float v = ld_sys_f32(p.mailbox[rank] + ((size_t)par * kArMaxW + r) * p.max_tensors + t);
        me = (v != v || me != me) ? NAN : fmaxf(me, v);
      }
      const int sh = aps_shift_from_exponent(me, fmt.exp_bits());
      s_shift[t] = sh;
      if (p.shifts_out && blockIdx.x == 0) p.shifts_out[t] = sh;
    }
    __syncthreads();
    // work item = a quarter chunk (1024 elements): one 4-element vector per thread, so every thread pays ONE peer
    // round trip (a whole chunk per CTA meant 4 dependent ~2 us round trips while most of the 296 CTAs sat idle)
    for (int w = blockIdx.x; w < 4 * p.n_chunks; w += gridDim.x) {
      const Chunk ck = p.chunks[w >> 2];
      const int sh = s_shift[ck.tensor];
      const float sc = pow2i(sh), inv = poison ? kPoison : pow2i(-sh);
      // all W peers' vectors requested before the first use
      for (int v4 = (w & 3) * kArThreads + threadIdx.x; v4 < (ck.len >> 2); v4 += 4 * kArThreads) {
        const long long e0 = ck.start + 4 * v4;
        // The fold below indexes the peers' vectors with a run-time rank, so `in` lives in local memory.  Loading straight
        // into it made the compiler emit LDG -> STL pairs through ONE register (the asm loads clobber memory, so each
        // local store had to precede the next load): 8 SERIALISED ~2 us peer round trips (fold 16 us at W = 8 against 3.7 us
        // at W = 2, profiles/allreduce_r2t2_*).  Land all W loads in distinct registers first, spill afterwards.
        uint4 in[kArMaxW], landed[kArMaxW];
#pragma unroll
        for (int r = 0; r < kArMaxW; ++r)
          if (r < W) landed[r] = p.weak_loads ? *reinterpret_cast<const uint4*>(p.data[r] + par * p.cap_bytes + e0 * 4)
                                              : ld_sys_u4(p.data[r] + par * p.cap_bytes + e0 * 4);
#pragma unroll
        for (int r = 0; r < kArMaxW; ++r)
          if (r < W) in[r] = landed[r];
        float s4[4] = {0.f, 0.f, 0.f, 0.f}, c4[4] = {0.f, 0.f, 0.f, 0.f};
        // rolled on purpose (the peers' vectors sit in a small local array): unrolling over 8 ranks only multiplies the code
#pragma unroll 1
        for (int r = 0; r < W; ++r) {
          {
            float v[4] = {__uint_as_float(in[r].x) * sc, __uint_as_float(in[r].y) * sc, __uint_as_float(in[r].z) * sc,
                          __uint_as_float(in[r].w) * sc};
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = fmt.q_wide(v[i]);
            if (KAHAN) fold_kahan_members_n(fmt, s4, c4, v);
            else fold_plain_members_n(fmt, s4, v);
          }
        }
        *reinterpret_cast<float4*>(p.grad + e0) = make_float4(s4[0] * inv, s4[1] * inv, s4[2] * inv, s4[3] * inv);
      }
    }
    stamp(p, 7);
    if (blockIdx.x == 0 && threadIdx.x == 0) p.ws[2] = epoch;
    return;
  }

  // ---------------- in-switch fp32 sum (ALGO_NVLS_SUM): the (8,23) path ---------------------------------------
  // Reference: plain NCCL all_reduce per parameter (dist_util.py:55-59).  Here: every rank publishes its fp32
  // gradients in the symmetric buffer; the owner of a chunk issues ONE multimem.ld_reduce.add.v4.f32 per 16 bytes -
  // the NVSwitch adds the W replicas - and ONE multimem.st broadcasts the sum into every rank's buffer; each rank
  // then copies its buffer back.  2(W-1)/W * 4 B per element cross the links, the SMs never fold.  All ranks receive
  // the same bits (one reducer per element); the summation order is the switch's, as NCCL's is NCCL's.
  if (WIRE == 32 && p.algo == ALGO_NVLS_SUM) {
    float* my_raw = reinterpret_cast<float*>(my_data);
    for (int c = blockIdx.x; c < p.n_chunks; c += gridDim.x) {
      const Chunk ck = p.chunks[c];
      for (int v4 = threadIdx.x; v4 < (ck.len >> 2); v4 += kArThreads)
        st_f4(my_raw + ck.start + 4 * v4, ld_stream_f4(p.grad + ck.start + 4 * v4));
    }
    grid_barrier(bar);
    stamp(p, 3);
    signal_peers(p, 1, epoch);
    poison |= wait_peers(p, 1, epoch);
    stamp(p, 4);
    float* mc = reinterpret_cast<float*>(p.mc_data + par * p.cap_bytes);
    for (int c = rank + W * (int)blockIdx.x; c < p.n_chunks; c += W * (int)gridDim.x) {
      const Chunk ck = p.chunks[c];
      for (int v4 = threadIdx.x; v4 < (ck.len >> 2); v4 += kArThreads) {
        float4 r = multimem_ld_reduce_add_f4(mc + ck.start + 4 * v4);
        if (poison) r = make_float4(kPoison, kPoison, kPoison, kPoison);
        multimem_st_f4(mc + ck.start + 4 * v4, r);
      }
    }
    stamp(p, 5);
    grid_barrier(bar);
    signal_peers(p, 2, epoch);
    poison |= wait_peers(p, 2, epoch);
    stamp(p, 6);
    for (int c = blockIdx.x; c < p.n_chunks; c += gridDim.x) {
      const Chunk ck = p.chunks[c];
      for (int v4 = threadIdx.x; v4 < (ck.len >> 2); v4 += kArThreads) {
        const uint4 u = ld_sys_u4(my_raw + ck.start + 4 * v4);
        float4 r = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
        if (poison) r = make_float4(kPoison, kPoison, kPoison, kPoison);
        st_f4(p.grad + ck.start + 4 * v4, r);
      }
    }
    stamp(p, 7);
    if (blockIdx.x == 0 && threadIdx.x == 0) p.ws[2] = epoch;
    return;
  }

  // ---------------- phase A: per-tensor amax, exponent exchange ---------------------
  if (p.use_aps) {
    // contiguous chunk range per CTA: consecutive chunks mostly belong to the same tensor, so the
    // per-thread maxima are carried across chunks and the block reduction + atomic happen only
    // when the tensor changes (large tensors: ~1 reduction per CTA instead of 1 per chunk)
    {
      const int per = (p.n_chunks + gridDim.x - 1) / gridDim.x;
      const int c0 = blockIdx.x * per, c1 = min(p.n_chunks, c0 + per);
      unsigned m = 0u;
      int cur_tensor = -1;
      for (int c = c0; c < c1; c += 2) {
        const bool has2 = c + 1 < c1;
        const Chunk ck = p.chunks[c];
        const Chunk ckn = has2 ? p.chunks[c + 1] : ck;
        // both chunks' vectors are requested before either is reduced
        uint4 x[4], y[4];
        const bool a0 = (int)threadIdx.x < (ck.len >> 4), a1 = has2 && (int)threadIdx.x < (ckn.len >> 4);
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        const uint4* q0 = reinterpret_cast<const uint4*>(p.grad + ck.start + 16 * threadIdx.x);
        const uint4* q1 = reinterpret_cast<const uint4*>(p.grad + ckn.start + 16 * threadIdx.x);
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = a0 ? q0[i] : z;
#pragma unroll
        for (int i = 0; i < 4; ++i) y[i] = a1 ? q1[i] : z;
        if (ck.tensor != cur_tensor) {
          if (cur_tensor >= 0) {
            m = block_max_u32(m, red);
            if (threadIdx.x == 0 && m) atomicMax(&amax[cur_tensor], m);
          }
          cur_tensor = ck.tensor;
          m = 0u;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
          m = max(max(m, x[i].x & 0x7FFFFFFFu), max(x[i].y & 0x7FFFFFFFu, max(x[i].z & 0x7FFFFFFFu, x[i].w & 0x7FFFFFFFu)));
        if (has2) {
          if (ckn.tensor != cur_tensor) {
            m = block_max_u32(m, red);
            if (threadIdx.x == 0 && m) atomicMax(&amax[cur_tensor], m);
            cur_tensor = ckn.tensor;
            m = 0u;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i)
            m = max(max(m, y[i].x & 0x7FFFFFFFu), max(y[i].y & 0x7FFFFFFFu, max(y[i].z & 0x7FFFFFFFu, y[i].w & 0x7FFFFFFFu)));
        }
      }
      if (cur_tensor >= 0) {
        m = block_max_u32(m, red);
        if (threadIdx.x == 0 && m) atomicMax(&amax[cur_tensor], m);
      }
    }
    grid_barrier(bar);
    stamp(p, 1);
    if (blockIdx.x == 0) {
      // PUSH the exponents into every rank's mailbox (slot [parity][source rank][tensor]) so that after the flag
      // round each rank reads only LOCAL memory: one NVLink round trip less on the critical path
      unsigned* amax_next = p.ws + 4 + (par ^ 1u) * p.max_tensors;
      for (int t = threadIdx.x; t < p.T; t += kArThreads) {
        const unsigned ab = amax[t];
        const float me = ab >= 0x7F800000u ? INFINITY : aps_local_exponent(__uint_as_float(ab), (float)W);
        for (int r = 0; r < W; ++r)
          asm volatile("st.relaxed.sys.global.f32 [%0], %1;" :: "l"(p.mailbox[r] + ((size_t)par * kArMaxW + rank) * p.max_tensors + t), "f"(me) : "memory");
      }
      // the other parity's amax buffer is used by the NEXT launch (possibly a different
      // bucket with more tensors): clear all of it
      for (int t = threadIdx.x; t < p.max_tensors; t += kArThreads) amax_next[t] = 0u;
      __syncthreads();
    }
    signal_peers(p, 0, epoch);
    poison |= wait_peers(p, 0, epoch);
    for (int t = threadIdx.x; t < p.T; t += kArThreads) {
      float me = -INFINITY;
      for (int r = 0; r < W; ++r) {
        const float v = ld_sys_f32(p.mailbox[rank] + ((size_t)par * kArMaxW + r) * p.max_tensors + t);
        me = (v != v || me != me) ? NAN : fmaxf(me, v);
      }
      const int sh = aps_shift_from_exponent(me, fmt.exp_bits());
      s_shift[t] = sh;
      if (p.shifts_out && blockIdx.x == 0) p.shifts_out[t] = sh;
      // guard counters: tensors whose global max is Inf/NaN, and all-zero tensors, since the first launch
      if (blockIdx.x == 0 && (me != me || me == INFINITY)) atomicAdd(p.ws + 4 + 2 * p.max_tensors + 16, 1u);
      if (blockIdx.x == 0 && me == -INFINITY) atomicAdd(p.ws + 4 + 2 * p.max_tensors + 17, 1u);
    }
    __syncthreads();
  }

  stamp(p, 2);
  // ---------------- phase B: scale, cast, pack into the local wire buffer -----------
  // two chunks per CTA iteration: 8 x 16 B loads in flight per thread before the first use (the local
  // passes are latency-bound: bytes in flight / ~1.2 us of loaded HBM latency)
  for (int c = blockIdx.x; c < p.n_chunks; c += 2 * gridDim.x) {
    const int c2 = c + gridDim.x;
    const bool has2 = c2 < p.n_chunks;
    const Chunk ck0 = p.chunks[c];
    const Chunk ck1 = has2 ? p.chunks[c2] : ck0;
    const bool a0 = (int)threadIdx.x < (ck0.len >> 4);
    const bool a1 = has2 && (int)threadIdx.x < (ck1.len >> 4);
    const long long e0 = ck0.start + 16 * threadIdx.x, e1 = ck1.start + 16 * threadIdx.x;
    float x0[16], x1[16];
    if (a0) load16(p.grad + e0, x0);
    if (a1) load16(p.grad + e1, x1);
    if (a0) pack_granule<WIRE>(p, fmt, x0, p.use_aps ? pow2i(s_shift[ck0.tensor]) : 1.f, my_data, e0);
    if (a1) pack_granule<WIRE>(p, fmt, x1, p.use_aps ? pow2i(s_shift[ck1.tensor]) : 1.f, my_data, e1);
  }
  grid_barrier(bar);
  stamp(p, 3);
  signal_peers(p, 1, epoch);
  poison |= wait_peers(p, 1, epoch);
  stamp(p, 4);

  // ---------------- phase C: rank-ordered quantised fold over peer memory ----------
  const bool two_shot = p.algo != ALGO_ONESHOT;
  // opt-in, NOT the CPD fold: the switch sums the W ranks' fp8 codes with an fp16 accumulator and rounds once
  // (ALGO_NVLS_FAST); everything around it (APS exponents, pack, gather, unscale) is the exact path's
  const bool nvls_fast = WIRE == 8 && FoldOps<F>::kHw && p.algo == ALGO_NVLS_FAST;
  const int c_begin = two_shot ? rank + W * (int)blockIdx.x : (int)blockIdx.x;
  const int c_step = two_shot ? W * (int)gridDim.x : (int)gridDim.x;
  for (int c = c_begin; c < p.n_chunks; c += c_step) {
    const Chunk ck = p.chunks[c];
    const float inv = poison ? kPoison : (p.use_aps ? pow2i(-s_shift[ck.tensor]) : 1.f);
    for (int g = threadIdx.x; g < (ck.len >> 4); g += kArThreads) {
      const long long e0 = ck.start + 16 * g;
      const bool first = c == c_begin && g == (int)threadIdx.x;      // finer stamps of CTA 0's first granule (slots 9..12)
      if (first) stamp(p, 9);
      // narrow wires: issue all W peer loads before the first use (W x 16..32 B in flight
      // per thread hides the ~2 us NVLink latency); fp32 wire loads rank by rank.
      constexpr int kPre = WIRE <= 16 ? kArMaxW : 1;
      Granule<WIRE> in[kPre];
      float s[16], cc[16];
      if (nvls_fast) {
        Granule<WIRE> sum;
        sum.v[0] = multimem_ld_reduce_add_fp8x16<IsE5M2<F>::value>(p.mc_data + par * p.cap_bytes + e0);
        granule_decode<WIRE>(fmt, sum, s);
      } else {
      if (WIRE <= 16) {
        // all W peer loads land in distinct registers BEFORE anything is spilled to the run-time indexed `in` (see the
        // raw variant above: loading into `in` directly serialised the W NVLink round trips)
        Granule<WIRE> landed[kPre];
#pragma unroll
        for (int r = 0; r < kPre; ++r)
          if (r < W) landed[r] = p.weak_loads ? granule_load_weak<WIRE>(p.data[r] + par * p.cap_bytes, e0)
                                              : granule_load_sys<WIRE>(p.data[r] + par * p.cap_bytes, e0);
#pragma unroll
        for (int r = 0; r < kPre; ++r)
          if (r < W) in[r] = landed[r];
      }
      if (first && WIRE <= 16) {
        unsigned acc = 0u;
#pragma unroll
        for (int r = 0; r < kPre; ++r)
          if (r < W) acc ^= in[r].v[0].x;
        asm volatile("" :: "r"(acc));                                   // all peer vectors have landed
        stamp(p, 10);
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) { s[i] = 0.f; cc[i] = 0.f; }
      if (WIRE == 32) {
        // fp32 wire (no APS / plain fp32): the granule goes in four 16-byte columns; for each, all W peers' vectors are
        // requested before the first is folded (the rank-by-rank version paid W serialised NVLink round trips:
        // plain fp32 256 KB on 8 GPUs 36.7 us against NCCL's 27.9)
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint4 landed[kArMaxW];
#pragma unroll
          for (int r = 0; r < kArMaxW; ++r)
            if (r < W) landed[r] = ld_sys_u4(p.data[r] + par * p.cap_bytes + e0 * 4 + 16 * q4);
#pragma unroll
          for (int r = 0; r < kArMaxW; ++r) {
            if (r < W) {
              const float v4[4] = {__uint_as_float(landed[r].x), __uint_as_float(landed[r].y), __uint_as_float(landed[r].z),
                                   __uint_as_float(landed[r].w)};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                if (KAHAN) fold_kahan(fmt, s[4 * q4 + i], cc[4 * q4 + i], v4[i]);
                else if (p.ident) s[4 * q4 + i] += v4[i];
                else fold_plain(fmt, s[4 * q4 + i], v4[i]);
              }
            }
          }
        }
      } else
      // the rank loop is not unrolled: the loads above are all in flight already, and 8 unrolled copies of decode +
      // 16 casts (with their cold paths) only multiply the SASS (194 KB -> ~90 KB rolled) without adding parallelism
#pragma unroll 1
      for (int r = 0; r < W; ++r) {
        {
          float v[16];
          if (WIRE <= 16) granule_decode<WIRE>(fmt, in[r % kPre], v);
          else granule_decode<WIRE>(fmt, granule_load_sys<WIRE>(p.data[r] + par * p.cap_bytes, e0), v);
          // WIRE < 32: every operand is a member of the format -> hardware-assisted exact casts, 16 behind one guard
          if (WIRE < 32) {
            if (KAHAN) fold_kahan_members_n(fmt, s, cc, v);
            else fold_plain_members_n(fmt, s, v);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (KAHAN) fold_kahan(fmt, s[i], cc[i], v[i]);
              else if (p.ident) s[i] += v[i];
              else fold_plain(fmt, s[i], v[i]);
            }
          }
        }
      }
      }  // !nvls_fast
      if (first) stamp(p, 11);
      if (poison) {                                     // also what travels to the peers in the gather half
#pragma unroll
        for (int i = 0; i < 16; ++i) s[i] = kPoison;
      }
      if (two_shot) {
        const Granule<WIRE> out = granule_encode<WIRE>(fmt, s);
        if ((p.algo == ALGO_NVLS || p.algo == ALGO_NVLS_FAST) && p.mc_gather) {
          granule_store_mc<WIRE>(p.mc_gather + par * p.cap_bytes, e0, out);
        } else {
#pragma unroll
          for (int r = 0; r < kArMaxW; ++r)
            if (r < W && r != rank) granule_store<WIRE>(p.gather[r] + par * p.cap_bytes, e0, out, true);
        }
      }
      if (p.use_aps || poison) {
#pragma unroll
        for (int i = 0; i < 16; ++i) s[i] *= inv;
      }
      store16(p.grad + e0, s);
      if (first) stamp(p, 12);
    }
  }

  // ---------------- phase D (two-shot): adopt the slices reduced by the peers ---------
  stamp(p, 5);
  if (two_shot) {
    grid_barrier(bar);
    signal_peers(p, 2, epoch);
    poison |= wait_peers(p, 2, epoch);
    stamp(p, 6);
    const unsigned char* my_gather = p.gather[rank] + par * p.cap_bytes;
    for (int c = blockIdx.x; c < p.n_chunks; c += 2 * gridDim.x) {
      const int c2 = c + gridDim.x;
      const bool u0 = (c % W) != rank, u1 = c2 < p.n_chunks && (c2 % W) != rank;
      const Chunk ck0 = p.chunks[c];
      const Chunk ck1 = u1 ? p.chunks[c2] : ck0;
      const bool a0 = u0 && (int)threadIdx.x < (ck0.len >> 4), a1 = u1 && (int)threadIdx.x < (ck1.len >> 4);
      const long long e0 = ck0.start + 16 * threadIdx.x, e1 = ck1.start + 16 * threadIdx.x;
      Granule<WIRE> in0, in1;
      if (a0) in0 = granule_load_sys<WIRE>(my_gather, e0);
      if (a1) in1 = granule_load_sys<WIRE>(my_gather, e1);
      if (a0) {
        const float inv = poison ? kPoison : (p.use_aps ? pow2i(-s_shift[ck0.tensor]) : 1.f);
        float v[16];
        granule_decode<WIRE>(fmt, in0, v);
        if (p.use_aps || poison) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] *= inv;
        }
        store16(p.grad + e0, v);
      }
      if (a1) {
        const float inv = poison ? kPoison : (p.use_aps ? pow2i(-s_shift[ck1.tensor]) : 1.f);
        float v[16];
        granule_decode<WIRE>(fmt, in1, v);
        if (p.use_aps || poison) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] *= inv;
        }
        store16(p.grad + e1, v);
      }
    }
  }

  stamp(p, 7);
  if (blockIdx.x == 0 && threadIdx.x == 0) p.ws[2] = epoch;
}

void allreduce_fused_cuda(const CommBuffers& cb, float* grad, const Chunk* chunks, int n_chunks, int T,
                          int e, int m, int rounding, bool use_aps, bool kahan, int algo, bool prequant, int wire,
                          unsigned* local_ws, int* shifts_out, int max_ctas, unsigned long long timeout_cycles,
                          cudaStream_t st) {
  if (n_chunks <= 0) return;
  ArParams p;
  p.rank = cb.rank; p.world = cb.world;
  for (int r = 0; r < kArMaxW; ++r) {
    p.data[r] = r < cb.world ? cb.data[r] : nullptr;
    p.gather[r] = r < cb.world ? cb.gather[r] : nullptr;
    p.mailbox[r] = r < cb.world ? cb.mailbox[r] : nullptr;
    p.flags[r] = r < cb.world ? cb.flags[r] : nullptr;
  }
  p.mc_gather = cb.mc_gather;
  p.mc_data = cb.mc_data;
  p.cap_bytes = (long long)cb.cap_bytes;
  p.max_tensors = cb.max_tensors;
  p.grad = grad; p.chunks = chunks; p.n_chunks = n_chunks; p.T = T;
  p.use_aps = use_aps; p.algo = algo & 0xFF; p.prequant = prequant;
  p.pack_hw = (algo >> 8) & 1;
  p.poll_cta0 = (algo >> 9) & 1;
  p.weak_loads = (algo >> 10) & 1;
  p.no_extra_fence = (algo >> 11) & 1;
  p.ident = (e == 8 && m == 23 && !kahan) ? 1 : 0;
  p.ws = local_ws; p.shifts_out = shifts_out; p.timeout = timeout_cycles;
  const bool fp8_hw = rounding == ROUND_REF && wire == 8 && ((e == 4 && m == 3) || (e == 5 && m == 2));
  if (p.algo == ALGO_NVLS_SUM && !(p.ident && !use_aps && wire == 32 && cb.mc_data)) p.algo = ALGO_TWOSHOT;
  if (p.algo == ALGO_NVLS_FAST && !(fp8_hw && use_aps && cb.mc_data && cb.mc_gather)) p.algo = ALGO_NVLS;
  if (p.algo == ALGO_NVLS && !cb.mc_gather) p.algo = ALGO_TWOSHOT;
  if (p.algo == ALGO_ONESHOT_RAW && !use_aps) p.algo = ALGO_ONESHOT;     // without APS the fp32-wire one-shot already needs one round
  // two co-resident CTAs per SM (<= 128 registers by launch bounds): 16 warps / SM keep enough 128-bit
  // loads in flight to cover HBM / NVLink latency; the grid barrier needs all CTAs resident.  A caller that
  // overlaps the reduction with other kernels passes a small max_ctas (8..32) so the comm grid co-runs with them.
  int grid = max_ctas > 0 ? max_ctas : device_sm_count(-1) * 2;
  // work items: whole chunks, except the single-round raw variant whose fold is scheduled in quarter chunks (one
  // 16-byte vector per thread = one peer round trip per thread)
  const int items = p.algo == ALGO_ONESHOT_RAW ? 4 * n_chunks : n_chunks;
  if (grid > items) grid = items;

#define CPD_AR_LAUNCH(WIRE_)                                                                          \
  do {                                                                                                \
    if (kahan) allreduce_fused_kernel<decltype(FMT), WIRE_, true><<<grid, kArThreads, 0, st>>>(p, FMT); \
    else allreduce_fused_kernel<decltype(FMT), WIRE_, false><<<grid, kArThreads, 0, st>>>(p, FMT);      \
  } while (0)

  // compile-time formats carry their natural wire width; everything else (and IEEE rounding) goes through the
  // runtime codec at the requested width
  const bool ref = rounding == ROUND_REF;
  if (ref && wire == 8 && e == 4 && m == 3) { FmtT<4, 3> FMT; CPD_AR_LAUNCH(8); }
  else if (ref && wire == 8 && e == 5 && m == 2) { FmtT<5, 2> FMT; CPD_AR_LAUNCH(8); }
  else if (ref && wire == 16 && e == 5 && m == 10) { FmtT<5, 10> FMT; CPD_AR_LAUNCH(16); }
  else if (ref && wire == 16 && e == 8 && m == 7) { FmtT<8, 7> FMT; CPD_AR_LAUNCH(16); }
  else {
    FmtR FMT{make_format(e, m, rounding)};
    if (wire == 8) CPD_AR_LAUNCH(8);
    else if (wire == 16) CPD_AR_LAUNCH(16);
    else CPD_AR_LAUNCH(32);
  }
#undef CPD_AR_LAUNCH
}

}  // namespace cpd
