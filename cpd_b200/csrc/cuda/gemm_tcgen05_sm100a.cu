// gemm_tcgen05_sm100a.cu — custom-precision GEMM on the 5th-gen tensor cores.
//
//   C[M,N] = A[M,K] . Bt[N,K]^T,  accumulator held at (exp, man):
//   for every K-slice:  partial = sum_{k in slice} a*b   (tcgen05.mma, fp32 in TMEM)
//                       s = Q(s (+) Q(partial))          (plain or Kahan, in registers)
//   Q = IEEE round-to-nearest-even cast to (exp, man) (rounding="ieee" of float_quantize); for
//   (5,10)/(8,7)/(4,3)/(5,2) it is two hardware convert instructions (fpformat_hw.cuh).
//
// This is the B200 replacement for the reference's tvm_gemm (float_kernel.cu:103-340): a
// 64-thread SIMT kernel with ~5 software casts per multiply-add.  Here the multiply-adds run on
// tcgen05 (operands staged by TMA into 128B-swizzled shared memory, accumulators in TMEM) and
// only ONE cast pair per output element per K-slice runs on the CUDA cores, overlapped with
// the next slice's MMAs through a double-buffered TMEM accumulator.
//
// Warp roles (640 threads, 1 CTA / SM, one 128x128 output tile per CTA):
//   warp 0      TMA producer   : cp.async.bulk.tensor 2D loads of A and B tiles, 4-stage mbarrier ring
//   warp 1      MMA issuer     : one elected lane issues tcgen05.mma (M128 x N128 x 32 B of K per
//                                instruction), tcgen05.commit frees smem stages / publishes TMEM stages
//   warp 2      TMEM allocator : 256 columns = 2 accumulator stages x 128 fp32 columns
//   warps 4-19  requant epilogue: four warpgroups, each owning 32 of the 128 columns; thread = one
//                                accumulator row (TMEM lane); tcgen05.ld 32x32b.x32 -> Q -> fold into
//                                the running (s, c) registers; final fp32 store (+ bias).  The requant
//                                is the ALU-bound part of the kernel, hence 512 of the 640 threads.
// Operand kinds: bf16 (kind::f16), fp32-as-tf32 (kind::tf32), e4m3 / e5m2 (kind::f8f6f4).
// M/N/K tails: TMA zero-fills out-of-bounds boxes, stores are predicated.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "../api.h"
#include "../common/fpformat.h"
#include "../common/fpformat_hw.cuh"
#include "format_dispatch.cuh"

namespace cpd {

namespace tc {

constexpr int BLOCK_M = 128, BLOCK_N = 128;
constexpr int STAGE_ROW_BYTES = 128;                      // one 128B swizzle row of K per stage
constexpr int A_STAGE_BYTES = BLOCK_M * STAGE_ROW_BYTES;  // 16 KB
constexpr int B_STAGE_BYTES = BLOCK_N * STAGE_ROW_BYTES;  // 16 KB
constexpr int NUM_STAGES = 4;
constexpr int NUM_ACC = 2;                                // TMEM accumulator stages
constexpr int TMEM_COLS = NUM_ACC * BLOCK_N;              // 256
constexpr int NUM_THREADS = 640;
constexpr int EPI_WARP0 = 4;
constexpr int EPI_THREADS = 512;
constexpr int EPI_COLS = 32;                               // columns per epilogue thread
constexpr int SMEM_BYTES = NUM_STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 /*align slack*/ + 256 /*barriers*/;

// ---- PTX wrappers ----------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :: "r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" :: "l"(tmap) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(smem_dst)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] . B[smem];  KIND: 0 f16/bf16, 1 tf32, 2 f8f6f4
template <int KIND>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (KIND == 0) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  } else if (KIND == 1) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  }
}
// arrive on an mbarrier once all previously issued MMAs have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row groups of 1024 B.
//   start addr >> 4 | LBO = 1 (unused for swizzled K-major) | SBO = 1024 B >> 4 | version 1 | SWIZZLE_128B
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct Barriers {
  uint64_t full[NUM_STAGES];
  uint64_t empty[NUM_STAGES];
  uint64_t tmem_full[NUM_ACC];
  uint64_t tmem_empty[NUM_ACC];
  uint32_t tmem_base;
};

}  // namespace tc

template <int KIND, class F, bool KAHAN>
__global__ void __launch_bounds__(tc::NUM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    float* __restrict__ C, const float* __restrict__ bias, int M, int N, int K, int elem_bytes,
                    int kblocks_per_slice, uint32_t idesc, F fmt) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + NUM_STAGES * A_STAGE_BYTES;
  Barriers* bars = reinterpret_cast<Barriers*>(smem + NUM_STAGES * (A_STAGE_BYTES + B_STAGE_BYTES));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BLOCK_M, n0 = blockIdx.x * BLOCK_N;
  const int k_elems_per_block = STAGE_ROW_BYTES / elem_bytes;
  const int num_kblocks = (K + k_elems_per_block - 1) / k_elems_per_block;
  const int num_slices = (num_kblocks + kblocks_per_slice - 1) / kblocks_per_slice;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NUM_STAGES; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
    for (int i = 0; i < NUM_ACC; ++i) { mbar_init(&bars->tmem_full[i], 1); mbar_init(&bars->tmem_empty[i], EPI_THREADS / 32); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_base, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int kb = 0; kb < num_kblocks; ++kb) {
        const int s = kb % NUM_STAGES;
        const uint32_t ph = (kb / NUM_STAGES) & 1;
        mbar_wait(&bars->empty[s], ph ^ 1);
        mbar_expect_tx(&bars->full[s], A_STAGE_BYTES + B_STAGE_BYTES);
        tma_load_2d(smem_a + s * A_STAGE_BYTES, &tmap_a, &bars->full[s], kb * k_elems_per_block, m0);
        tma_load_2d(smem_b + s * B_STAGE_BYTES, &tmap_b, &bars->full[s], kb * k_elems_per_block, n0);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int kb = 0;
      for (int sl = 0; sl < num_slices; ++sl) {
        const int acc = sl % NUM_ACC;
        const uint32_t acc_ph = (sl / NUM_ACC) & 1;
        mbar_wait(&bars->tmem_empty[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_addr = tmem_base + acc * BLOCK_N;
        const int kb_end = min(num_kblocks, kb + kblocks_per_slice);
        bool first = true;
        for (; kb < kb_end; ++kb) {
          const int s = kb % NUM_STAGES;
          const uint32_t ph = (kb / NUM_STAGES) & 1;
          mbar_wait(&bars->full[s], ph);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc(smem_u32(smem_a + s * A_STAGE_BYTES));
          const uint64_t bdesc = make_smem_desc(smem_u32(smem_b + s * B_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < STAGE_ROW_BYTES / 32; ++k) {          // 32 bytes of K per instruction
            umma<KIND>(d_addr, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, first ? 0u : 1u);
            first = false;
          }
          umma_commit(&bars->empty[s]);                              // smem stage reusable once these MMAs retire
        }
        umma_commit(&bars->tmem_full[acc]);                          // slice accumulator complete
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== requant epilogue =====================
    const int ew = warp - EPI_WARP0;           // 0..15
    const int quarter = warp & 3;              // TMEM lane quarter this warp may touch
    const int cg = ew >> 2;                    // which 32 columns
    const int row = quarter * 32 + lane;       // accumulator row == TMEM lane
    float s[EPI_COLS], c[KAHAN ? EPI_COLS : 1];
#pragma unroll
    for (int j = 0; j < EPI_COLS; ++j) s[j] = 0.f;
    if (KAHAN) {
#pragma unroll
      for (int j = 0; j < EPI_COLS; ++j) c[KAHAN ? j : 0] = 0.f;
    }
    for (int sl = 0; sl < num_slices; ++sl) {
      const int acc = sl % NUM_ACC;
      const uint32_t acc_ph = (sl / NUM_ACC) & 1;
      mbar_wait(&bars->tmem_full[acc], acc_ph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BLOCK_N + cg * EPI_COLS;
      uint32_t v[32];
      tmem_ld_32x32(taddr, v);
      tmem_ld_wait();
      // the TMEM stage is free as soon as the values are in registers: release it before the casts
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);            // 16 warp arrivals free the TMEM stage
#pragma unroll
      for (int j = 0; j < EPI_COLS; ++j) {
        const float p = fmt.q(__uint_as_float(v[j]));
        if (KAHAN) fold_kahan(fmt, s[j], c[KAHAN ? j : 0], p);
        else fold_plain(fmt, s[j], p);
      }
    }
    // final store (+ bias): thread owns 32 contiguous floats of one row
    const int gm = m0 + row;
    const int gn = n0 + cg * EPI_COLS;
    if (gm < M && gn < N) {
      float* crow = C + (long long)gm * N + gn;
      const int ncols = min(EPI_COLS, N - gn);
      if (ncols == EPI_COLS && (N & 3) == 0) {
#pragma unroll
        for (int j = 0; j < EPI_COLS; j += 4) {
          float4 o = make_float4(s[j], s[j + 1], s[j + 2], s[j + 3]);
          if (bias) {
            const float4 bb = *reinterpret_cast<const float4*>(bias + gn + j);
            o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
          }
          *reinterpret_cast<float4*>(crow + j) = o;
        }
      } else {
#pragma unroll
        for (int j = 0; j < EPI_COLS; ++j)
          if (j < ncols) crow[j] = s[j] + (bias ? bias[gn + j] : 0.f);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- host side -----------------------------------------------------------------------
namespace {

using PFN_encodeTiled = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess) fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 2-D row-major [rows, K] operand, box = {128 B of K, 128 rows}, 128B swizzle, zero fill out of bounds
bool make_tmap(CUtensorMap* tm, const void* base, int rows, int K, int elem_bytes, CUtensorMapDataType dt) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)K * elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)(tc::STAGE_ROW_BYTES / elem_bytes), (cuuint32_t)tc::BLOCK_M};
  cuuint32_t estr[2] = {1, 1};
  return enc(tm, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int KIND>
int launch_kind(const CUtensorMap& ta, const CUtensorMap& tb, float* C, const float* bias, int M, int N, int K, int elem_bytes,
                int kblocks_per_slice, uint32_t idesc, int e, int m, bool kahan, cudaStream_t st) {
  dim3 grid((N + tc::BLOCK_N - 1) / tc::BLOCK_N, (M + tc::BLOCK_M - 1) / tc::BLOCK_M);
  int rc = 0;
  // accumulator casts use IEEE rounding (hardware F2F/F2FP converts for fp16/bf16/e4m3/e5m2-shaped formats)
  CPD_DISPATCH_FORMAT_IEEE_HW(e, m, FMT, {
    if (kahan) {
      auto kfn = gemm_tcgen05_kernel<KIND, decltype(FMT), true>;
      if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES) != cudaSuccess) rc = -3;
      else kfn<<<grid, tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(ta, tb, C, bias, M, N, K, elem_bytes, kblocks_per_slice, idesc, FMT);
    } else {
      auto kfn = gemm_tcgen05_kernel<KIND, decltype(FMT), false>;
      if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES) != cudaSuccess) rc = -3;
      else kfn<<<grid, tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(ta, tb, C, bias, M, N, K, elem_bytes, kblocks_per_slice, idesc, FMT);
    }
  });
  return rc;
}

}  // namespace

// operand_mode: 0 bf16 (kind::f16), 1 fp32 read as tf32 (kind::tf32), 2 e4m3, 3 e5m2 (kind::f8f6f4)
int gemm_tc_cuda(const void* A, const void* Bt, float* C, int M, int N, int K, int e, int m, bool kahan, int k_slice,
                 int operand_mode, const float* bias, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  int elem_bytes, kind;
  uint32_t fmt_code;
  CUtensorMapDataType dt;
  switch (operand_mode) {
    case 0: elem_bytes = 2; kind = 0; fmt_code = 1; dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; break;
    case 1: elem_bytes = 4; kind = 1; fmt_code = 2; dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; break;
    case 2: elem_bytes = 1; kind = 2; fmt_code = 0; dt = CU_TENSOR_MAP_DATA_TYPE_UINT8; break;
    case 3: elem_bytes = 1; kind = 2; fmt_code = 1; dt = CU_TENSOR_MAP_DATA_TYPE_UINT8; break;
    default: return -1;
  }
  if (((long long)K * elem_bytes) % 16 != 0) return -2;       // TMA global stride must be a multiple of 16 B
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(Bt)) & 15u) return -2;
  const int k_elems_per_block = tc::STAGE_ROW_BYTES / elem_bytes;
  int kblocks_per_slice = k_slice <= 0 ? (K + k_elems_per_block - 1) / k_elems_per_block
                                       : (k_slice + k_elems_per_block - 1) / k_elems_per_block;
  if (kblocks_per_slice < 1) kblocks_per_slice = 1;
  CUtensorMap ta, tb;
  if (!make_tmap(&ta, A, M, K, elem_bytes, dt) || !make_tmap(&tb, Bt, N, K, elem_bytes, dt)) return -4;
  // instruction descriptor: D fp32 | A/B format | K-major both | N >> 3 | M >> 4
  const uint32_t idesc = (1u << 4) | (fmt_code << 7) | (fmt_code << 10) | ((uint32_t)(tc::BLOCK_N >> 3) << 17) |
                         ((uint32_t)(tc::BLOCK_M >> 4) << 24);
  if (kind == 0) return launch_kind<0>(ta, tb, C, bias, M, N, K, elem_bytes, kblocks_per_slice, idesc, e, m, kahan, st);
  if (kind == 1) return launch_kind<1>(ta, tb, C, bias, M, N, K, elem_bytes, kblocks_per_slice, idesc, e, m, kahan, st);
  return launch_kind<2>(ta, tb, C, bias, M, N, K, elem_bytes, kblocks_per_slice, idesc, e, m, kahan, st);
}

}  // namespace cpd
