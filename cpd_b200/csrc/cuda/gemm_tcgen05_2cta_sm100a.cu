// gemm_tcgen05_2cta_sm100a.cu — the custom-precision GEMM on CTA PAIRS (thread-block clusters).
//
// Same semantics as gemm_tcgen05_sm100a.cu (C = A . Bt^T, accumulator re-quantised to (exp,man)
// with IEEE rounding after every K-slice), restructured around what bounds the 1-CTA kernel:
//   * operand traffic: a cluster of 2 CTAs on one TPC issues tcgen05.mma.cta_group::2 with
//     M = 256, N = 256; each CTA stages only its own 128 rows of A and its own 128 rows of Bt, so
//     every byte fetched from L2 feeds twice as many MACs as in the 128x128 kernel (64 vs 128 B/clk/SM);
//   * requant cost: for the formats that fit in 16 bits — (5,10) = fp16, (8,7) = bf16, (4,3)/(5,2) =
//     fp8 values held as fp16 — the running accumulator is literally KEPT at that precision in
//     packed 16-bit registers: one F2FP.PACK_AB converts (= casts) two values at once, the
//     64 columns a thread owns take 32 registers (64 with the Kahan compensation), and the
//     requant is ~4 instructions per element per slice instead of ~25.
// Pipeline per CTA (640 threads): warp 0 TMA producer (5-stage ring; both CTAs credit the leader's
// mbarrier), warp 1 of the LEADER issues the MMAs for the pair and multicasts the commits, warp 2
// allocates all 512 TMEM columns (2 accumulator stages x 256 fp32 columns) with cta_group::2,
// warps 4-19 are the requant epilogue (thread = one of the CTA's 128 rows x 64 columns).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "../api.h"
#include "tcgen05_utils.cuh"

namespace cpd {

namespace tc2 {
using namespace tc;
constexpr int CTA_M = 128, PAIR_M = 256, BLOCK_N = 256, CTA_BROWS = 128;
constexpr int A_STAGE_BYTES = CTA_M * STAGE_ROW_BYTES;      // 16 KB
constexpr int B_STAGE_BYTES = CTA_BROWS * STAGE_ROW_BYTES;  // 16 KB
constexpr int NUM_STAGES = 5;
constexpr int NUM_ACC = 2;
constexpr int TMEM_COLS = NUM_ACC * BLOCK_N;                // 512: all of TMEM
constexpr int NUM_THREADS = 640;
constexpr int EPI_WARP0 = 4, EPI_WARPS = 16;
constexpr int EPI_COLS = 64;                                // columns per epilogue thread
constexpr int SMEM_BYTES = NUM_STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 + 256;

struct Barriers {
  uint64_t full[NUM_STAGES];
  uint64_t empty[NUM_STAGES];
  uint64_t tmem_full[NUM_ACC];
  uint64_t tmem_empty[NUM_ACC];
  uint32_t tmem_base;
};

// ---- accumulator storage policies: two values per 32-bit register, cast == convert ------------------
struct AccF16 {        // (5,10): the accumulator IS an fp16 pair
  static __device__ __forceinline__ uint32_t q2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  static __device__ __forceinline__ float2 up(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }
};
struct AccBF16 {       // (8,7)
  static __device__ __forceinline__ uint32_t q2(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  static __device__ __forceinline__ float2 up(uint32_t v) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v)); }
};
template <__nv_fp8_interpretation_t INTERP, int LIMIT>   // (4,3)/(5,2): fp8 value set, held exactly as fp16
struct AccFp8 {
  static __device__ __forceinline__ uint32_t q2(float a, float b) {
    const __nv_fp8x2_storage_t c = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, INTERP);
    const __half2_raw hr = __nv_cvt_fp8x2_to_halfraw2(c, INTERP);
    uint32_t r = (uint32_t)hr.x | ((uint32_t)hr.y << 16);
    if (fmaxf(fabsf(a), fabsf(b)) >= (float)LIMIT) {        // CPD reserves the top exponent: overflow -> Inf (rare)
      if (fabsf(a) >= (float)LIMIT) r = (r & 0xFFFF0000u) | (a < 0.f ? 0xFC00u : 0x7C00u);
      if (fabsf(b) >= (float)LIMIT) r = (r & 0x0000FFFFu) | (b < 0.f ? 0xFC000000u : 0x7C000000u);
    }
    return r;
  }
  static __device__ __forceinline__ float2 up(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }
};
using AccE4M3 = AccFp8<__NV_E4M3, 248>;
using AccE5M2 = AccFp8<__NV_E5M2, 61440>;

}  // namespace tc2

// ACC: packed-accumulator policy; single_slice kernels skip the running accumulator entirely.
template <int KIND, class ACC, bool KAHAN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(tc2::NUM_THREADS, 1)
gemm_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         float* __restrict__ C, const float* __restrict__ bias, int M, int N, int K, int elem_bytes,
                         int kblocks_per_slice, uint32_t idesc, int tiles_m, int tiles_n, int num_tiles, tc::ConvA conv) {
  using namespace tc2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + NUM_STAGES * A_STAGE_BYTES;
  Barriers* bars = reinterpret_cast<Barriers*>(smem + NUM_STAGES * (A_STAGE_BYTES + B_STAGE_BYTES));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();           // 0 = leader (issues the MMAs)
  // Persistent: CTA pair p walks output tiles p, p + #pairs, ... with RUNNING smem-stage / TMEM-stage counters, so the
  // prologue (barriers, TMEM allocation, cluster syncs, pipeline fill) is paid once per pair and tile i's final requant +
  // store overlaps tile i+1's first MMAs (the TMEM stage is handed back as soon as it is in registers).
  const int pair0 = blockIdx.x >> 1, pair_step = gridDim.x >> 1;
#define CPD_TILE2_COORDS(t)                                                                          \
  const int band_ = (t) / (tc::RASTER_GROUP * tiles_n), bm0_ = band_ * tc::RASTER_GROUP;             \
  const int bsz_ = min(tc::RASTER_GROUP, tiles_m - bm0_), bi_ = (t) - band_ * (tc::RASTER_GROUP * tiles_n); \
  const int tile_m = bm0_ + bi_ % bsz_, tile_n = bi_ / bsz_;    /* banded raster, see the 1-CTA kernel */ \
  const int m0 = tile_m * PAIR_M + (int)cta_rank * CTA_M;          /* this CTA's rows of A / C */    \
  const int n0 = tile_n * BLOCK_N;                                                                   \
  const int nb0 = n0 + (int)cta_rank * CTA_BROWS;                  /* this CTA's rows of Bt */       \
  (void)m0; (void)n0; (void)nb0;
  const int k_elems_per_block = STAGE_ROW_BYTES / elem_bytes;
  const int num_kblocks = (K + k_elems_per_block - 1) / k_elems_per_block;
  const int num_slices = (num_kblocks + kblocks_per_slice - 1) / kblocks_per_slice;

  if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_a); prefetch_tmap(&tmap_b); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NUM_STAGES; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
    for (int i = 0; i < NUM_ACC; ++i) { mbar_init(&bars->tmem_full[i], 1); mbar_init(&bars->tmem_empty[i], 2 * EPI_WARPS); }
    fence_barrier_init();
  }
  cluster_sync_all();                                    // both CTAs' barriers exist before anyone signals remotely
  if (warp == 2) tmem_alloc_2cta(&bars->tmem_base, TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      uint32_t it = 0;                                     // running K-block count across tiles
      for (int t = pair0; t < num_tiles; t += pair_step) {
        CPD_TILE2_COORDS(t)
        int cw = 0, ch = 0, cn = 0;                        // implicit GEMM: base pixel of this CTA's first output position
        if (conv.cblocks) {
          cn = m0 / conv.OHW;
          const int rem = m0 - cn * conv.OHW, p = rem / conv.OW;
          ch = p * conv.stride - conv.pad;
          cw = (rem - p * conv.OW) * conv.stride - conv.pad;
        }
        for (int kb = 0; kb < num_kblocks; ++kb, ++it) {
          const int s = it % NUM_STAGES;
          const uint32_t ph = (it / NUM_STAGES) & 1;
          mbar_wait(&bars->empty[s], ph ^ 1);
          if (cta_rank == 0) mbar_expect_tx(&bars->full[s], 2 * (A_STAGE_BYTES + B_STAGE_BYTES));   // bytes of BOTH CTAs
          if (conv.cblocks) {
            const int tap = kb / conv.cblocks, cb = kb - tap * conv.cblocks;
            const int r = tap / conv.KW, sx = tap - r * conv.KW;
            tma_load_im2col_4d_2cta(smem_a + s * A_STAGE_BYTES, &tmap_a, &bars->full[s], cb * k_elems_per_block, cw, ch, cn,
                                    (uint16_t)(sx * conv.dil), (uint16_t)(r * conv.dil));
          } else
          tma_load_2d_2cta(smem_a + s * A_STAGE_BYTES, &tmap_a, &bars->full[s], kb * k_elems_per_block, m0);
          tma_load_2d_2cta(smem_b + s * B_STAGE_BYTES, &tmap_b, &bars->full[s], kb * k_elems_per_block, nb0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (cta_rank == 0 && lane == 0) {
      uint32_t it = 0, sl_run = 0;                         // running K-block / K-slice counts across tiles
      for (int t = pair0; t < num_tiles; t += pair_step) {
      int kb = 0;
      for (int sl = 0; sl < num_slices; ++sl, ++sl_run) {
        const int acc = sl_run % NUM_ACC;
        const uint32_t acc_ph = (sl_run / NUM_ACC) & 1;
        mbar_wait(&bars->tmem_empty[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_addr = tmem_base + acc * BLOCK_N;
        const int kb_end = min(num_kblocks, kb + kblocks_per_slice);
        bool first = true;
        for (; kb < kb_end; ++kb, ++it) {
          const int s = it % NUM_STAGES;
          const uint32_t ph = (it / NUM_STAGES) & 1;
          mbar_wait(&bars->full[s], ph);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc(smem_u32(smem_a + s * A_STAGE_BYTES));
          const uint64_t bdesc = make_smem_desc(smem_u32(smem_b + s * B_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < STAGE_ROW_BYTES / 32; ++k) {
            umma_2cta<KIND>(d_addr, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, first ? 0u : 1u);
            first = false;
          }
          umma_commit_2cta(&bars->empty[s]);             // frees this stage in BOTH CTAs
        }
        umma_commit_2cta(&bars->tmem_full[acc]);         // slice accumulator complete, in BOTH CTAs
      }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== requant epilogue =====================
    const int ew = warp - EPI_WARP0;
    const int quarter = warp & 3;
    const int cg = ew >> 2;                              // 64-column group
    const int row = quarter * 32 + lane;
    const uint32_t leader_tmem_empty0 = mapa_shared(smem_u32(&bars->tmem_empty[0]), 0);
    const uint32_t leader_tmem_empty1 = mapa_shared(smem_u32(&bars->tmem_empty[1]), 0);
    uint32_t sl_run = 0;                                 // running K-slice count across tiles (TMEM stage / phase)
    for (int t = pair0; t < num_tiles; t += pair_step) {
    CPD_TILE2_COORDS(t)
    uint32_t s2[EPI_COLS / 2], c2[KAHAN ? EPI_COLS / 2 : 1];
#pragma unroll
    for (int j = 0; j < EPI_COLS / 2; ++j) s2[j] = 0u;
    if (KAHAN) {
#pragma unroll
      for (int j = 0; j < EPI_COLS / 2; ++j) c2[KAHAN ? j : 0] = 0u;
    }
    for (int sl = 0; sl < num_slices; ++sl, ++sl_run) {
      const int acc = sl_run % NUM_ACC;
      const uint32_t acc_ph = (sl_run / NUM_ACC) & 1;
      mbar_wait(&bars->tmem_full[acc], acc_ph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BLOCK_N + cg * EPI_COLS;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + h * 32, v);
        tmem_ld_wait();
        if (h == 1) {                                    // both halves are in registers: hand the TMEM stage back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(acc == 0 ? leader_tmem_empty0 : leader_tmem_empty1);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int idx = h * 16 + j;
          const uint32_t p2 = ACC::q2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));      // Q(partial)
          const float2 p = ACC::up(p2);
          const float2 s = ACC::up(s2[idx]);
          if (KAHAN) {
            const float2 c = ACC::up(c2[KAHAN ? idx : 0]);
            const float2 y = ACC::up(ACC::q2(p.x - c.x, p.y - c.y));
            const uint32_t t2 = ACC::q2(s.x + y.x, s.y + y.y);
            const float2 t = ACC::up(t2);
            const float2 d = ACC::up(ACC::q2(t.x - s.x, t.y - s.y));
            c2[KAHAN ? idx : 0] = ACC::q2(d.x - y.x, d.y - y.y);
            s2[idx] = t2;
          } else {
            s2[idx] = ACC::q2(s.x + p.x, s.y + p.y);
          }
        }
      }
    }
    const int gm = m0 + row;
    const int gn = n0 + cg * EPI_COLS;
    if (gm < M && gn < N) {
      float* crow = C + (long long)gm * N + gn;
      const int ncols = min(EPI_COLS, N - gn);
      if (ncols == EPI_COLS && (N & 3) == 0) {
#pragma unroll
        for (int j = 0; j < EPI_COLS / 4; ++j) {
          const float2 a = ACC::up(s2[2 * j]), b = ACC::up(s2[2 * j + 1]);
          float4 o = make_float4(a.x, a.y, b.x, b.y);
          if (bias) {
            const float4 bb = *reinterpret_cast<const float4*>(bias + gn + 4 * j);
            o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
          }
          *reinterpret_cast<float4*>(crow + 4 * j) = o;
        }
      } else {
#pragma unroll
        for (int j = 0; j < EPI_COLS / 2; ++j) {
          const float2 a = ACC::up(s2[j]);
          if (2 * j < ncols) crow[2 * j] = a.x + (bias ? bias[gn + 2 * j] : 0.f);
          if (2 * j + 1 < ncols) crow[2 * j + 1] = a.y + (bias ? bias[gn + 2 * j + 1] : 0.f);
        }
      }
    }
    }
  }
#undef CPD_TILE2_COORDS

  tc_fence_before();
  cluster_sync_all();                                    // the peer may still be reading our smem / signalling our barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, TMEM_COLS);
  }
}

// ---- host side ---------------------------------------------------------------------------------
// gemm_tcgen05_sm100a.cu
bool make_tmap_im2col(CUtensorMap* tm, const void* base, const ConvGeom& g, int elem_bytes, CUtensorMapDataType dt, int box_pixels);
bool conv_geom_ok(const ConvGeom& g, int elem_bytes, int M, int K);
tc::ConvA conv_device_geom(const ConvGeom* g, int elem_bytes);

namespace {

using PFN_encodeTiled = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool make_tmap2(CUtensorMap* tm, const void* base, int rows, int K, long long ld, int elem_bytes, CUtensorMapDataType dt, int box_rows) {
  static PFN_encodeTiled enc = nullptr;
  if (!enc) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return false;
    enc = reinterpret_cast<PFN_encodeTiled>(p);
  }
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)(tc::STAGE_ROW_BYTES / elem_bytes), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(tm, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int KIND, class ACC>
int launch2(const CUtensorMap& ta, const CUtensorMap& tb, float* C, const float* bias, int M, int N, int K, int elem_bytes,
            int kblocks_per_slice, uint32_t idesc, bool kahan, cudaStream_t st, const tc::ConvA& conv) {
  const int tiles_m = (M + tc2::PAIR_M - 1) / tc2::PAIR_M, tiles_n = (N + tc2::BLOCK_N - 1) / tc2::BLOCK_N;
  const int num_tiles = tiles_m * tiles_n;
  static int max_pairs = 0;                                // co-resident CTA pairs (clusters of 2; a GPC with an odd SM count loses one)
  if (!max_pairs) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    max_pairs = sms > 1 ? sms / 2 : 1;
    auto probe = gemm_tcgen05_2cta_kernel<KIND, ACC, false>;
    if (cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::SMEM_BYTES) == cudaSuccess) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * max_pairs); cfg.blockDim = dim3(tc2::NUM_THREADS); cfg.dynamicSmemBytes = tc2::SMEM_BYTES;
      cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
      cfg.attrs = &at; cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, probe, &cfg) == cudaSuccess && n > 0 && n < max_pairs) max_pairs = n;
    }
    (void)cudaGetLastError();
  }
  const dim3 grid(2 * (num_tiles < max_pairs ? num_tiles : max_pairs));
  if (kahan) {
    auto kfn = gemm_tcgen05_2cta_kernel<KIND, ACC, true>;
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::SMEM_BYTES) != cudaSuccess) return -3;
    kfn<<<grid, tc2::NUM_THREADS, tc2::SMEM_BYTES, st>>>(ta, tb, C, bias, M, N, K, elem_bytes, kblocks_per_slice, idesc, tiles_m, tiles_n, num_tiles, conv);
  } else {
    auto kfn = gemm_tcgen05_2cta_kernel<KIND, ACC, false>;
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::SMEM_BYTES) != cudaSuccess) return -3;
    kfn<<<grid, tc2::NUM_THREADS, tc2::SMEM_BYTES, st>>>(ta, tb, C, bias, M, N, K, elem_bytes, kblocks_per_slice, idesc, tiles_m, tiles_n, num_tiles, conv);
  }
  return 0;
}

template <int KIND>
int launch2_fmt(const CUtensorMap& ta, const CUtensorMap& tb, float* C, const float* bias, int M, int N, int K, int elem_bytes,
                int kbs, uint32_t idesc, int e, int m, bool kahan, cudaStream_t st, const tc::ConvA& conv) {
  if (e == 5 && m == 10) return launch2<KIND, tc2::AccF16>(ta, tb, C, bias, M, N, K, elem_bytes, kbs, idesc, kahan, st, conv);
  if (e == 8 && m == 7) return launch2<KIND, tc2::AccBF16>(ta, tb, C, bias, M, N, K, elem_bytes, kbs, idesc, kahan, st, conv);
  if (e == 4 && m == 3) return launch2<KIND, tc2::AccE4M3>(ta, tb, C, bias, M, N, K, elem_bytes, kbs, idesc, kahan, st, conv);
  if (e == 5 && m == 2) return launch2<KIND, tc2::AccE5M2>(ta, tb, C, bias, M, N, K, elem_bytes, kbs, idesc, kahan, st, conv);
  return -5;
}

}  // namespace

// Returns 0 on launch, <0 if this kernel does not cover the request (the caller then uses the 1-CTA kernel).
int gemm_tc2_cuda(const void* A, const void* Bt, float* C, int M, int N, int K, int e, int m, bool kahan, int k_slice,
                  int operand_mode, const float* bias, cudaStream_t st, long long lda, long long ldb, const ConvGeom* conv) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (lda <= 0) lda = K;
  if (ldb <= 0) ldb = K;
  const bool packed = (e == 5 && m == 10) || (e == 8 && m == 7) || (e == 4 && m == 3) || (e == 5 && m == 2);
  if (!packed) return -5;
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
  if (conv) {
    if (!conv_geom_ok(*conv, elem_bytes, M, K)) return -7;
    lda = K;
  }
  if ((lda * elem_bytes) % 16 != 0 || (ldb * elem_bytes) % 16 != 0) return -2;
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(Bt)) & 15u) return -2;
  const int k_elems_per_block = tc::STAGE_ROW_BYTES / elem_bytes;
  int kbs = k_slice <= 0 ? (K + k_elems_per_block - 1) / k_elems_per_block : (k_slice + k_elems_per_block - 1) / k_elems_per_block;
  if (kbs < 1) kbs = 1;
  CUtensorMap ta, tb;
  if (!(conv ? make_tmap_im2col(&ta, A, *conv, elem_bytes, dt, tc2::CTA_M) : make_tmap2(&ta, A, M, K, lda, elem_bytes, dt, tc2::CTA_M)) ||
      !make_tmap2(&tb, Bt, N, K, ldb, elem_bytes, dt, tc2::CTA_BROWS)) return -4;
  const tc::ConvA cgeom = conv_device_geom(conv, elem_bytes);
  const uint32_t idesc = (1u << 4) | (fmt_code << 7) | (fmt_code << 10) | ((uint32_t)(tc2::BLOCK_N >> 3) << 17) |
                         ((uint32_t)(tc2::PAIR_M >> 4) << 24);
  if (kind == 0) return launch2_fmt<0>(ta, tb, C, bias, M, N, K, elem_bytes, kbs, idesc, e, m, kahan, st, cgeom);
  if (kind == 1) return launch2_fmt<1>(ta, tb, C, bias, M, N, K, elem_bytes, kbs, idesc, e, m, kahan, st, cgeom);
  return launch2_fmt<2>(ta, tb, C, bias, M, N, K, elem_bytes, kbs, idesc, e, m, kahan, st, cgeom);
}

}  // namespace cpd
