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
// PERSISTENT: the grid is min(#tiles, #SMs) CTAs and every CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...
// (tile = (split, n-tile, m-tile), m fastest so that neighbouring CTAs share the B tile in L2).  Barrier initialisation,
// TMEM allocation and the tensor-map prefetch are paid once per CTA instead of once per tile - for the convolution shapes
// (K = 576: nine K-blocks per tile) that prologue used to cost more than the tile's MMAs - and the three roles keep
// RUNNING stage / phase counters across tiles, so the TMA ring and the two TMEM accumulator stages stay full over tile
// boundaries: the last slice's requant + store of tile i overlaps the first MMAs of tile i + 1.
// Warp roles (640 threads, 1 CTA / SM, 128x128 output tiles):
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
#include "tcgen05_utils.cuh"
#include "../common/fpformat.h"
#include "../common/fpformat_hw.cuh"
#include "format_dispatch.cuh"

namespace cpd {

namespace tc {

constexpr int BLOCK_M = 128, BLOCK_N = 128;
constexpr int A_STAGE_BYTES = BLOCK_M * STAGE_ROW_BYTES;  // 16 KB
constexpr int B_STAGE_BYTES = BLOCK_N * STAGE_ROW_BYTES;  // 16 KB
constexpr int NUM_STAGES = 4;                              // measured: 7 stages (all of shared memory) = +2.6 % with one requant at the end, -6 % with a requant every 256 (profiles/r2/t11_bench_gemm_7stages.json)
constexpr int NUM_ACC = 2;                                // TMEM accumulator stages
constexpr int TMEM_COLS = NUM_ACC * BLOCK_N;              // 256
constexpr int NUM_THREADS = 640;
constexpr int EPI_WARP0 = 4;
constexpr int EPI_THREADS = 512;
constexpr int EPI_COLS = 32;                               // columns per epilogue thread
constexpr int SMEM_BYTES = NUM_STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 /*align slack*/ + 256 /*barriers*/;

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
                    int kblocks_per_slice, uint32_t idesc, F fmt, int kblocks_per_split, int tiles_m, int tiles_n, int splits,
                    tc::ConvA conv) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + NUM_STAGES * A_STAGE_BYTES;
  Barriers* bars = reinterpret_cast<Barriers*>(smem + NUM_STAGES * (A_STAGE_BYTES + B_STAGE_BYTES));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k_elems_per_block = STAGE_ROW_BYTES / elem_bytes;
  const int total_kblocks = (K + k_elems_per_block - 1) / k_elems_per_block;
  const int num_tiles = tiles_m * tiles_n * splits;
  float* const C_base = C;

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

  // tile -> (split z, n-tile, m-tile) and the split's K-block range
  // Raster: bands of RASTER_GROUP m-tiles, m fastest inside a band, then n, then the next band - the tiles in flight
  // at any moment (one per SM) form a ~8 x 18 block of the tile grid and share ~26 operand tiles instead of ~150
  // (m-fastest over the whole matrix re-read A once per n-tile: the 32768 x 256 x 2304 convolution GEMM lost 12 %).
#define CPD_TILE_COORDS(t)                                                                          \
  const int tz_ = (t) / (tiles_m * tiles_n), tr_ = (t) - tz_ * (tiles_m * tiles_n);                  \
  const int band_ = tr_ / (RASTER_GROUP * tiles_n), bm0_ = band_ * RASTER_GROUP;                     \
  const int bsz_ = min(RASTER_GROUP, tiles_m - bm0_), bi_ = tr_ - band_ * (RASTER_GROUP * tiles_n);  \
  const int tm_ = bm0_ + bi_ % bsz_, tn_ = bi_ / bsz_;                                              \
  const int m0 = tm_ * BLOCK_M, n0 = tn_ * BLOCK_N;                                                 \
  const int kb0 = tz_ * kblocks_per_split;                                                          \
  const int num_kblocks = max(0, min(total_kblocks - kb0, kblocks_per_split));                      \
  const int num_slices = (num_kblocks + kblocks_per_slice - 1) / kblocks_per_slice;                 \
  (void)m0; (void)n0; (void)kb0; (void)num_slices;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;                                   // running K-block count: smem stage = it % NUM_STAGES
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        CPD_TILE_COORDS(t)
        int cw = 0, ch = 0, cn = 0;                      // implicit GEMM: base pixel of the tile's first output position
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
          mbar_expect_tx(&bars->full[s], A_STAGE_BYTES + B_STAGE_BYTES);
          if (conv.wgrad) {
            const int pix0 = (kb0 + kb) * 64;                        // K-block = 64 output positions
            const int pn = pix0 / conv.OHW, rem = pix0 - pn * conv.OHW, pp = rem / conv.OW;
            const int ph_ = pp * conv.stride - conv.pad, pw_ = (rem - pp * conv.OW) * conv.stride - conv.pad;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              tma_load_2d(smem_a + s * A_STAGE_BYTES + j * (A_STAGE_BYTES / 2), &tmap_a, &bars->full[s], m0 + 64 * j, pix0);
              int tap = (n0 + 64 * j) / conv.C;
              const int c = (n0 + 64 * j) - tap * conv.C;
              tap = min(tap, conv.taps - 1);                         // columns past N are computed on a valid tap and never stored
              const int r = tap / conv.KW, sx = tap - r * conv.KW;
              tma_load_im2col_4d(smem_b + s * B_STAGE_BYTES + j * (B_STAGE_BYTES / 2), &tmap_b, &bars->full[s], c, pw_, ph_, pn,
                                 (uint16_t)(sx * conv.dil), (uint16_t)(r * conv.dil));
            }
            continue;
          }
          if (conv.cblocks) {
            const int tap = (kb0 + kb) / conv.cblocks, cb = (kb0 + kb) - tap * conv.cblocks;
            const int r = tap / conv.KW, sx = tap - r * conv.KW;
            tma_load_im2col_4d(smem_a + s * A_STAGE_BYTES, &tmap_a, &bars->full[s], cb * k_elems_per_block, cw, ch, cn,
                               (uint16_t)(sx * conv.dil), (uint16_t)(r * conv.dil));
          } else {
            tma_load_2d(smem_a + s * A_STAGE_BYTES, &tmap_a, &bars->full[s], (kb0 + kb) * k_elems_per_block, m0);
          }
          tma_load_2d(smem_b + s * B_STAGE_BYTES, &tmap_b, &bars->full[s], (kb0 + kb) * k_elems_per_block, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t it = 0, sl_run = 0;                       // running K-block / K-slice counts across tiles
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        CPD_TILE_COORDS(t)
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
            if (conv.wgrad) {                                          // MN-major tiles: 16 pixels (K) = 2048 B per instruction
              const uint64_t adesc = make_smem_desc_mn(smem_u32(smem_a + s * A_STAGE_BYTES), conv.lbo, conv.sbo);
              const uint64_t bdesc = make_smem_desc_mn(smem_u32(smem_b + s * B_STAGE_BYTES), conv.lbo, conv.sbo);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma<KIND>(d_addr, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc, first ? 0u : 1u);
                first = false;
              }
            } else {
            const uint64_t adesc = make_smem_desc(smem_u32(smem_a + s * A_STAGE_BYTES));
            const uint64_t bdesc = make_smem_desc(smem_u32(smem_b + s * B_STAGE_BYTES));
#pragma unroll
            for (int k = 0; k < STAGE_ROW_BYTES / 32; ++k) {          // 32 bytes of K per instruction
              umma<KIND>(d_addr, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, first ? 0u : 1u);
              first = false;
            }
            }
            umma_commit(&bars->empty[s]);                              // smem stage reusable once these MMAs retire
          }
          umma_commit(&bars->tmem_full[acc]);                          // slice accumulator complete
        }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== requant epilogue =====================
    const int ew = warp - EPI_WARP0;           // 0..15
    const int quarter = warp & 3;              // TMEM lane quarter this warp may touch
    const int cg = ew >> 2;                    // which 32 columns
    const int row = quarter * 32 + lane;       // accumulator row == TMEM lane
    uint32_t sl_run = 0;                       // running K-slice count across tiles (TMEM stage / phase)
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      CPD_TILE_COORDS(t)
      float* Ct = C_base + (long long)tz_ * M * N;                     // split-K: this split's partial matrix
      float s[EPI_COLS], c[KAHAN ? EPI_COLS : 1];
#pragma unroll
      for (int j = 0; j < EPI_COLS; ++j) s[j] = 0.f;
      if (KAHAN) {
#pragma unroll
        for (int j = 0; j < EPI_COLS; ++j) c[KAHAN ? j : 0] = 0.f;
      }
      for (int sl = 0; sl < num_slices; ++sl, ++sl_run) {
        const int acc = sl_run % NUM_ACC;
        const uint32_t acc_ph = (sl_run / NUM_ACC) & 1;
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
        float* crow = Ct + (long long)gm * N + gn;
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
  }
#undef CPD_TILE_COORDS

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
bool make_tmap(CUtensorMap* tm, const void* base, int rows, int K, long long ld, int elem_bytes, CUtensorMapDataType dt,
               int box_rows = tc::BLOCK_M) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};          // the TRUE K: the tail of the last box is zero-filled
  cuuint64_t gstride[1] = {(cuuint64_t)ld * elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)(tc::STAGE_ROW_BYTES / elem_bytes), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(tm, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// NHWC activation tensor read in im2col mode: one load = box_pixels consecutive output positions x (128 B of channels) of
// one filter tap; the bounding box [-pad, dim + pad - (K-1)*dil) and the traversal stride make the hardware enumerate
// exactly the OW x OH output positions of every image, the halo is zero-filled
bool make_tmap_im2col(CUtensorMap* tm, const void* base, const ConvGeom& g, int elem_bytes, CUtensorMapDataType dt, int box_pixels) {
  using PFN_encodeIm2col = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static PFN_encodeIm2col enc = nullptr;
  if (!enc) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return false;
    enc = reinterpret_cast<PFN_encodeIm2col>(p);
  }
  cuuint64_t gdim[4] = {(cuuint64_t)g.C, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.B};
  cuuint64_t gstride[3] = {(cuuint64_t)g.C * elem_bytes, (cuuint64_t)g.W * g.C * elem_bytes, (cuuint64_t)g.H * g.W * g.C * elem_bytes};
  int lower[2] = {-g.pad, -g.pad};
  int upper[2] = {g.pad - (g.KW - 1) * g.dil, g.pad - (g.KH - 1) * g.dil};
  cuuint32_t estr[4] = {1, (cuuint32_t)g.stride, (cuuint32_t)g.stride, 1};
  return enc(tm, dt, 4, const_cast<void*>(base), gdim, gstride, lower, upper, (cuuint32_t)(tc::STAGE_ROW_BYTES / elem_bytes),
             (cuuint32_t)box_pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// does the implicit-GEMM path cover this convolution?  (corner / offset ranges of the descriptor, whole channel blocks)
bool conv_geom_ok(const ConvGeom& g, int elem_bytes, int M, int K) {
  const int cpb = tc::STAGE_ROW_BYTES / elem_bytes;
  if (g.C <= 0 || g.C % cpb) return false;
  if (g.stride < 1 || g.stride > 8 || g.dil < 1 || g.pad < 0 || g.pad > 127) return false;
  if ((g.KH - 1) * g.dil > 255 || (g.KW - 1) * g.dil > 255) return false;
  if (g.pad - (g.KW - 1) * g.dil < -128 || g.pad - (g.KH - 1) * g.dil < -128) return false;
  if (g.OH != (g.H + 2 * g.pad - g.dil * (g.KH - 1) - 1) / g.stride + 1 || g.OW != (g.W + 2 * g.pad - g.dil * (g.KW - 1) - 1) / g.stride + 1) return false;
  if (g.OH < 1 || g.OW < 1) return false;
  return (long long)g.B * g.OH * g.OW == M && (long long)g.KH * g.KW * g.C == K;
}

tc::ConvA conv_device_geom(const ConvGeom* g, int elem_bytes) {
  tc::ConvA c = {0, 1, 1, 1, 0, 1, 1, 0, 1, 1, 8192, 1024};
  if (g) {
    c.cblocks = g->C / (tc::STAGE_ROW_BYTES / elem_bytes);
    c.OHW = g->OH * g->OW; c.OW = g->OW; c.stride = g->stride; c.pad = g->pad; c.dil = g->dil; c.KW = g->KW;
  }
  return c;
}

namespace {

// C = fold_z Q(C (+) part_z) (+ bias): ordered low-precision fold of the split-K partials (each already a member of
// the format: the splits' own last cast), same cast as the GEMM epilogue
template <class F, bool KAHAN>
__global__ void __launch_bounds__(256)
splitk_fold_kernel(const float* __restrict__ parts, int splits, long long MN, int N, const float* __restrict__ bias,
                   float* __restrict__ C, F fmt) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < MN; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f, c = 0.f;
    for (int z = 0; z < splits; ++z) {
      const float p = parts[(long long)z * MN + i];
      if (KAHAN) fold_kahan(fmt, s, c, p);
      else fold_plain(fmt, s, p);
    }
    C[i] = s + (bias ? bias[i % N] : 0.f);
  }
}

template <int KIND>
int launch_kind(const CUtensorMap& ta, const CUtensorMap& tb, float* C, const float* bias, int M, int N, int K, int elem_bytes,
                int kblocks_per_slice, uint32_t idesc, int e, int m, bool kahan, cudaStream_t st, int splits, int kblocks_per_split,
                float* workspace, const tc::ConvA& conv) {
  const int tiles_n = (N + tc::BLOCK_N - 1) / tc::BLOCK_N, tiles_m = (M + tc::BLOCK_M - 1) / tc::BLOCK_M;
  const long long num_tiles = (long long)tiles_m * tiles_n * splits;
  const int sms = device_sm_count(-1);
  dim3 grid((unsigned)(num_tiles < sms ? num_tiles : sms));               // persistent: one CTA per SM walks the tiles
  float* out = splits > 1 ? workspace : C;
  const float* kbias = splits > 1 ? nullptr : bias;
  int rc = 0;
  // accumulator casts use IEEE rounding (hardware F2F/F2FP converts for fp16/bf16/e4m3/e5m2-shaped formats)
  CPD_DISPATCH_FORMAT_IEEE_HW(e, m, FMT, {
    if (kahan) {
      auto kfn = gemm_tcgen05_kernel<KIND, decltype(FMT), true>;
      if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES) != cudaSuccess) rc = -3;
      else kfn<<<grid, tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(ta, tb, out, kbias, M, N, K, elem_bytes, kblocks_per_slice, idesc, FMT, kblocks_per_split, tiles_m, tiles_n, splits, conv);
    } else {
      auto kfn = gemm_tcgen05_kernel<KIND, decltype(FMT), false>;
      if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES) != cudaSuccess) rc = -3;
      else kfn<<<grid, tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(ta, tb, out, kbias, M, N, K, elem_bytes, kblocks_per_slice, idesc, FMT, kblocks_per_split, tiles_m, tiles_n, splits, conv);
    }
    if (rc == 0 && splits > 1) {
      const long long MN = (long long)M * N;
      const int fgrid = (int)((MN + 255) / 256 < 4096 ? (MN + 255) / 256 : 4096);
      if (kahan) splitk_fold_kernel<decltype(FMT), true><<<fgrid, 256, 0, st>>>(workspace, splits, MN, N, bias, C, FMT);
      else splitk_fold_kernel<decltype(FMT), false><<<fgrid, 256, 0, st>>>(workspace, splits, MN, N, bias, C, FMT);
    }
  });
  return rc;
}

}  // namespace

// operand_mode: 0 bf16 (kind::f16), 1 fp32 read as tf32 (kind::tf32), 2 e4m3, 3 e5m2 (kind::f8f6f4)
int gemm_tc_cuda(const void* A, const void* Bt, float* C, int M, int N, int K, int e, int m, bool kahan, int k_slice,
                 int operand_mode, const float* bias, int variant, cudaStream_t st, long long lda, long long ldb, int splits,
                 float* workspace, const ConvGeom* conv) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (lda <= 0) lda = K;
  if (ldb <= 0) ldb = K;
  if (splits < 1) splits = 1;
  if (splits > 1 && !workspace) return -6;
  // CTA-pair kernel when it covers the format and its 256x256 tiles still fill the machine
  // (Kahan's four casts per element make the kernel requant-bound either way; measured slightly faster on 128x128 tiles)
  if (splits == 1 && (variant == 2 || (variant == 0 && !kahan && 2 * ((M + 255) / 256) * ((N + 255) / 256) >= device_sm_count(-1)))) {
    const int rc = gemm_tc2_cuda(A, Bt, C, M, N, K, e, m, kahan, k_slice, operand_mode, bias, st, lda, ldb, conv);
    if (rc == 0 || variant == 2) return rc;
  }
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
  if ((lda * elem_bytes) % 16 != 0 || (ldb * elem_bytes) % 16 != 0) return -2;       // TMA global stride must be a multiple of 16 B
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(Bt)) & 15u) return -2;
  const int k_elems_per_block = tc::STAGE_ROW_BYTES / elem_bytes;
  const int total_kblocks = (K + k_elems_per_block - 1) / k_elems_per_block;
  int kblocks_per_slice = k_slice <= 0 ? total_kblocks : (k_slice + k_elems_per_block - 1) / k_elems_per_block;
  if (kblocks_per_slice < 1) kblocks_per_slice = 1;
  // every split owns whole K-slices, so the per-slice casts inside a split are those of the unsplit kernel
  int slices = (total_kblocks + kblocks_per_slice - 1) / kblocks_per_slice;
  if (splits > slices) splits = slices;
  const int slices_per_split = (slices + splits - 1) / splits;
  const int kblocks_per_split = slices_per_split * kblocks_per_slice;
  splits = (total_kblocks + kblocks_per_split - 1) / kblocks_per_split;
  CUtensorMap ta, tb;
  if (!(conv ? make_tmap_im2col(&ta, A, *conv, elem_bytes, dt, tc::BLOCK_M) : make_tmap(&ta, A, M, K, lda, elem_bytes, dt)) ||
      !make_tmap(&tb, Bt, N, K, ldb, elem_bytes, dt)) return -4;
  const tc::ConvA cgeom = conv_device_geom(conv, elem_bytes);
  // instruction descriptor: D fp32 | A/B format | K-major both | N >> 3 | M >> 4
  const uint32_t idesc = (1u << 4) | (fmt_code << 7) | (fmt_code << 10) | ((uint32_t)(tc::BLOCK_N >> 3) << 17) |
                         ((uint32_t)(tc::BLOCK_M >> 4) << 24);
  if (kind == 0) return launch_kind<0>(ta, tb, C, bias, M, N, K, elem_bytes, kblocks_per_slice, idesc, e, m, kahan, st, splits, kblocks_per_split, workspace, cgeom);
  if (kind == 1) return launch_kind<1>(ta, tb, C, bias, M, N, K, elem_bytes, kblocks_per_slice, idesc, e, m, kahan, st, splits, kblocks_per_split, workspace, cgeom);
  return launch_kind<2>(ta, tb, C, bias, M, N, K, elem_bytes, kblocks_per_slice, idesc, e, m, kahan, st, splits, kblocks_per_split, workspace, cgeom);
}


// Weight gradient of a convolution as an implicit GEMM: dW[O, KH*KW*C] = dy[pixels, O]^T . im2col(x)[pixels, KH*KW*C],
// reduction over pixels = B*OH*OW (split-K + ordered fold as in gemm_tc_cuda).  dY: [pixels, O] and X: [B, H, W, C], both
// dense bf16; neither a transposed dy nor a transposed im2col matrix is ever written (MN-major operand tiles).
int gemm_wgrad_cuda(const void* dY, const void* X, float* dW, const ConvGeom& g, int O, int e, int m, bool kahan, int k_slice,
                    cudaStream_t st, int splits, float* workspace, int lbo, int sbo) {
  const int elem_bytes = 2;
  const long long pixels = (long long)g.B * g.OH * g.OW, N = (long long)g.KH * g.KW * g.C;
  if (O <= 0 || pixels <= 0 || pixels > INT32_MAX || N > INT32_MAX) return -1;
  if (!conv_geom_ok(g, elem_bytes, (int)pixels, (int)N)) return -7;
  if ((O * elem_bytes) % 16 != 0) return -2;
  if ((reinterpret_cast<uintptr_t>(dY) | reinterpret_cast<uintptr_t>(X)) & 15u) return -2;
  if (splits < 1) splits = 1;
  if (splits > 1 && !workspace) return -6;
  const int M = O, K = (int)pixels;
  const int total_kblocks = (K + 63) / 64;
  int kblocks_per_slice = k_slice <= 0 ? total_kblocks : (k_slice + 63) / 64;
  if (kblocks_per_slice < 1) kblocks_per_slice = 1;
  const int slices = (total_kblocks + kblocks_per_slice - 1) / kblocks_per_slice;
  if (splits > slices) splits = slices;
  const int slices_per_split = (slices + splits - 1) / splits;
  const int kblocks_per_split = slices_per_split * kblocks_per_slice;
  splits = (total_kblocks + kblocks_per_split - 1) / kblocks_per_split;
  CUtensorMap ta, tb;
  // dy as a 2-D tensor [pixels rows, O contiguous]: box = {64 o (128 B), 64 pixels}
  if (!make_tmap(&ta, dY, K, O, O, elem_bytes, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 64) ||
      !make_tmap_im2col(&tb, X, g, elem_bytes, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 64)) return -4;
  tc::ConvA c = conv_device_geom(&g, elem_bytes);
  c.wgrad = 1; c.C = g.C; c.taps = g.KH * g.KW;
  c.lbo = lbo > 0 ? lbo : tc::A_STAGE_BYTES / 2;
  c.sbo = sbo > 0 ? sbo : 1024;
  // bf16 | A and B MN-major (bits 15, 16) | N >> 3 | M >> 4
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(tc::BLOCK_N >> 3) << 17) |
                         ((uint32_t)(tc::BLOCK_M >> 4) << 24);
  return launch_kind<0>(ta, tb, dW, nullptr, M, (int)N, K, elem_bytes, kblocks_per_slice, idesc, e, m, kahan, st, splits, kblocks_per_split,
                        workspace, c);
}

}  // namespace cpd
