// gemm_tcgen05_mx_sm100a.cu — block-scaled fp8 variant of the custom-precision GEMM.
//
// Operands are MXFP8: E4M3 (or E5M2) codes with one UE8M0 power-of-two scale per 32 consecutive
// K-elements of every row (produced by mx_quantize_sm100.cu).  The tensor core applies the scales
// itself: tcgen05.mma.kind::mxf8f6f4.block_scale reads the A/B scale factors from TMEM, where
// tcgen05.cp places the 512-byte scale-factor atoms that travel through the same mbarrier ring as
// the operand tiles.  Everything else (TMA ring, double-buffered TMEM accumulator, per-K-slice
// IEEE requant of the running accumulator in the epilogue warps) is the 1-CTA kernel's design.
// This is the "(4,3)/(5,2)" operand path: block scaling keeps fp8 operands accurate for
// activations/gradients whose magnitude varies along K — the same power-of-two-shift idea as APS.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "../api.h"
#include "../common/fpformat.h"
#include "../common/fpformat_hw.cuh"
#include "tcgen05_utils.cuh"

namespace cpd {

namespace tcmx {
using namespace tc;
constexpr int BLOCK_M = 128, BLOCK_N = 128;
constexpr int A_STAGE_BYTES = BLOCK_M * STAGE_ROW_BYTES;   // 16 KB: 128 rows x 128 fp8
constexpr int B_STAGE_BYTES = BLOCK_N * STAGE_ROW_BYTES;
constexpr int SF_BYTES = 512;                              // one scale-factor atom: 128 rows x 4 k-blocks
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES + 2 * SF_BYTES;
constexpr int NUM_STAGES = 4;
constexpr int NUM_ACC = 2;
constexpr int TMEM_COLS = 512;                             // 2 x 128 accumulator columns + scale-factor columns
constexpr int SFA_COL = 256, SFB_COL = 272;
constexpr int NUM_THREADS = 640, EPI_WARP0 = 4, EPI_WARPS = 16, EPI_COLS = 32;
constexpr int SMEM_BYTES = NUM_STAGES * STAGE_BYTES + 1024 + 256;

struct Barriers {
  uint64_t full[NUM_STAGES];
  uint64_t empty[NUM_STAGES];
  uint64_t tmem_full[NUM_ACC];
  uint64_t tmem_empty[NUM_ACC];
  uint32_t tmem_base;
};

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// 32 rows x 16 B, no swizzle: 8-row groups of 128 contiguous bytes
__device__ __forceinline__ uint64_t make_sf_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;              // LBO (unused: a single 16-byte column)
  d |= (uint64_t)(128 >> 4) << 32;     // SBO: 8-row group stride
  d |= (uint64_t)1 << 46;              // descriptor version (Blackwell)
  return d;                            // layout type 0 = no swizzle
}
__device__ __forceinline__ void tmem_cp_sf(uint32_t tmem_dst, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" :: "r"(tmem_dst), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void umma_mx(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t tmem_sfa,
                                        uint32_t tmem_sfb, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %6, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%4], [%5], p;\n\t}"
               :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(tmem_sfa), "r"(tmem_sfb), "r"(accumulate) : "memory");
}
}  // namespace tcmx

template <class F, bool KAHAN>
__global__ void __launch_bounds__(tcmx::NUM_THREADS, 1)
gemm_tcgen05_mx_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                       const uint8_t* __restrict__ sfa, const uint8_t* __restrict__ sfb, float* __restrict__ C,
                       const float* __restrict__ bias, int M, int N, int Kp, int kblocks_per_slice, uint32_t idesc_base, F fmt) {
  using namespace tcmx;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + NUM_STAGES * A_STAGE_BYTES;
  uint8_t* smem_sf = smem + NUM_STAGES * (A_STAGE_BYTES + B_STAGE_BYTES);          // [stage][A 512 | B 512]
  Barriers* bars = reinterpret_cast<Barriers*>(smem + NUM_STAGES * STAGE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BLOCK_M, n0 = blockIdx.x * BLOCK_N;
  const int num_kblocks = Kp >> 7;
  const int num_slices = (num_kblocks + kblocks_per_slice - 1) / kblocks_per_slice;

  if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_a); prefetch_tmap(&tmap_b); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NUM_STAGES; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
    for (int i = 0; i < NUM_ACC; ++i) { mbar_init(&bars->tmem_full[i], 1); mbar_init(&bars->tmem_empty[i], EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_base, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kblocks; ++kb) {
        const int s = kb % NUM_STAGES;
        const uint32_t ph = (kb / NUM_STAGES) & 1;
        mbar_wait(&bars->empty[s], ph ^ 1);
        mbar_expect_tx(&bars->full[s], STAGE_BYTES);
        tma_load_2d(smem_a + s * A_STAGE_BYTES, &tmap_a, &bars->full[s], kb * 128, m0);
        tma_load_2d(smem_b + s * B_STAGE_BYTES, &tmap_b, &bars->full[s], kb * 128, n0);
        bulk_load(smem_sf + s * 2 * SF_BYTES, sfa + ((long long)blockIdx.y * num_kblocks + kb) * SF_BYTES, SF_BYTES, &bars->full[s]);
        bulk_load(smem_sf + s * 2 * SF_BYTES + SF_BYTES, sfb + ((long long)blockIdx.x * num_kblocks + kb) * SF_BYTES, SF_BYTES, &bars->full[s]);
      }
    }
  } else if (warp == 1) {
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
          // scale factors: smem atom -> TMEM (replicated to the four lane quarters); in-order with the MMAs below
          tmem_cp_sf(tmem_base + SFA_COL, make_sf_desc(smem_u32(smem_sf + s * 2 * SF_BYTES)));
          tmem_cp_sf(tmem_base + SFB_COL, make_sf_desc(smem_u32(smem_sf + s * 2 * SF_BYTES + SF_BYTES)));
          const uint64_t adesc = make_smem_desc(smem_u32(smem_a + s * A_STAGE_BYTES));
          const uint64_t bdesc = make_smem_desc(smem_u32(smem_b + s * B_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k) {            // 32 fp8 elements (= one scale block) of K per instruction
            const uint32_t idesc = idesc_base | ((uint32_t)k << 4) | ((uint32_t)k << 29);       // b_sf_id | a_sf_id = k
            umma_mx(d_addr, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, tmem_base + SFA_COL, tmem_base + SFB_COL,
                    first ? 0u : 1u);
            first = false;
          }
          umma_commit(&bars->empty[s]);
        }
        umma_commit(&bars->tmem_full[acc]);
      }
    }
  } else if (warp >= EPI_WARP0) {
    const int ew = warp - EPI_WARP0;
    const int quarter = warp & 3;
    const int cg = ew >> 2;
    const int row = quarter * 32 + lane;
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
#pragma unroll
      for (int j = 0; j < EPI_COLS; ++j) {
        const float p = fmt.q(__uint_as_float(v[j]));
        if (KAHAN) fold_kahan(fmt, s[j], c[KAHAN ? j : 0], p);
        else fold_plain(fmt, s[j], p);
      }
    }
    const int gm = m0 + row, gn = n0 + cg * EPI_COLS;
    if (gm < M && gn < N) {
      float* crow = C + (long long)gm * N + gn;
      const int ncols = min(EPI_COLS, N - gn);
#pragma unroll
      for (int j = 0; j < EPI_COLS; ++j)
        if (j < ncols) crow[j] = s[j] + (bias ? bias[gn + j] : 0.f);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

namespace {
using PFN_encodeTiled = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
bool make_tmap_u8(CUtensorMap* tm, const void* base, int rows, int Kp) {
  static PFN_encodeTiled enc = nullptr;
  if (!enc) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return false;
    enc = reinterpret_cast<PFN_encodeTiled>(p);
  }
  cuuint64_t gdim[2] = {(cuuint64_t)Kp, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)Kp};
  cuuint32_t box[2] = {128, 128};
  cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

// codes_a [M, Kp], codes_b [N, Kp] (Kp % 128 == 0), sf_* in atom layout (mx_quantize_cuda)
int gemm_tc_mx_cuda(const uint8_t* codes_a, const uint8_t* sf_a, const uint8_t* codes_b, const uint8_t* sf_b, float* C, int M, int N,
                    int Kp, int e, int m, bool kahan, int k_slice, bool e5m2, const float* bias, cudaStream_t st) {
  if (M <= 0 || N <= 0 || Kp <= 0) return 0;
  if (Kp % 128) return -2;
  CUtensorMap ta, tb;
  if (!make_tmap_u8(&ta, codes_a, M, Kp) || !make_tmap_u8(&tb, codes_b, N, Kp)) return -4;
  int kbs = k_slice <= 0 ? (Kp >> 7) : (k_slice + 127) / 128;
  if (kbs < 1) kbs = 1;
  const uint32_t f = e5m2 ? 1u : 0u;
  // block-scaled instruction descriptor: formats | K-major | N>>3 | scale format UE8M0 (bit 23) | M>>4; sf ids are OR-ed in per MMA
  const uint32_t idesc = (f << 7) | (f << 10) | ((uint32_t)(tcmx::BLOCK_N >> 3) << 17) | (1u << 23) | ((uint32_t)(tcmx::BLOCK_M >> 4) << 24);
  dim3 grid((N + tcmx::BLOCK_N - 1) / tcmx::BLOCK_N, (M + tcmx::BLOCK_M - 1) / tcmx::BLOCK_M);
  int rc = 0;
  CPD_DISPATCH_FORMAT_IEEE_HW(e, m, FMT, {
    if (kahan) {
      auto kfn = gemm_tcgen05_mx_kernel<decltype(FMT), true>;
      if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, tcmx::SMEM_BYTES) != cudaSuccess) rc = -3;
      else kfn<<<grid, tcmx::NUM_THREADS, tcmx::SMEM_BYTES, st>>>(ta, tb, sf_a, sf_b, C, bias, M, N, Kp, kbs, idesc, FMT);
    } else {
      auto kfn = gemm_tcgen05_mx_kernel<decltype(FMT), false>;
      if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, tcmx::SMEM_BYTES) != cudaSuccess) rc = -3;
      else kfn<<<grid, tcmx::NUM_THREADS, tcmx::SMEM_BYTES, st>>>(ta, tb, sf_a, sf_b, C, bias, M, N, Kp, kbs, idesc, FMT);
    }
  });
  return rc;
}

}  // namespace cpd
