// api.h — torch-free C++ entry points of the native library (CPU ops, CUDA
// launchers, comm runtime).  bindings.cpp is the only file that sees PyTorch.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#ifndef CPD_NO_CUDA
#include <cuda_runtime.h>
#else
typedef void* cudaStream_t;
#endif

namespace cpd {

// One schedulable piece of a flat multi-tensor arena: `len` elements starting at
// `start`, all inside tensor `tensor` (so one APS shift / one trust ratio applies).
struct Chunk {
  long long start;
  int len;
  int tensor;
};

// ------------------------------------------------------------------ CPU ops --
void quantize_cpu(const float* in, float* out, int64_t n, int e, int m, int rounding);
void quantize_scaled_cpu(const float* in, float* out, int64_t n, int e, int m, int rounding, int shift);
// C[M,N] = A[M,K] * B[K,N]; per-MAC quantised (optionally Kahan) sequential-k accumulation
// A(i,k) = A[i*sa0 + k*sa1], B(k,j) = B[k*sb0 + j*sb1]  (strides express transposes)
void gemm_exact_cpu(const float* A, int64_t sa0, int64_t sa1, const float* B, int64_t sb0, int64_t sb1,
                    float* C, int64_t ldc, int M, int N, int K, int e, int m, int rounding, bool kahan);
// tile semantics: fp32 partial over each K-slice (operands optionally pre-rounded to
// an operand format), then s = Q(s (+) Q(partial)) per slice.
void gemm_tile_cpu(const float* A, int64_t sa0, int64_t sa1, const float* B, int64_t sb0, int64_t sb1,
                   float* C, int64_t ldc, int M, int N, int K, int e, int m, int rounding, bool kahan,
                   int k_slice, int op_e, int op_m, int splits = 1);
// APS + rank-ordered (Kahan) fold over W source buffers that share one segment table.
void aps_fold_cpu(const float* const* srcs, int W, float* out, const int64_t* seg_off, int T,
                  int e, int m, int rounding, bool use_aps, bool kahan, float world_mult,
                  int* shifts_out);
// LARS / SGD multi-tensor step on fp32 masters (reference mix.py:297-314 semantics).
void lars_step_cpu(float* master, float* model, const float* grad, float* mom, const int64_t* seg_off,
                   int T, float lr, float momentum, float wd, float coeff, float eps, bool use_lars,
                   bool nesterov, const float* wd_mult);

// ----------------------------------------------------------------- CUDA ops --
#ifndef CPD_NO_CUDA
int device_sm_count(int dev);

void quantize_cuda(const float* in, float* out, int64_t n, int e, int m, int rounding, int shift,
                   cudaStream_t st);
// fused amax + quantize statistics (optional): writes |x| max bits into *amax_bits
void amax_cuda(const float* in, int64_t n, unsigned* amax_bits, cudaStream_t st);

void gemm_exact_cuda(const float* A, int64_t sa0, int64_t sa1, const float* B, int64_t sb0, int64_t sb1,
                     float* C, int64_t ldc, int M, int N, int K, int e, int m, int rounding, bool kahan,
                     cudaStream_t st);

// Multi-source APS fold (local sources: emulate_node slots, or W==1): one launch, up to kMaxLocalSources arenas.
constexpr int kMaxLocalSources = 64;
//   srcs: device array of W pointers to flat fp32 arenas sharing the chunk table
//   workspace: 4 + 2*T_cap unsigned (grid barrier, launch counter, amax[2][T_cap]), zeroed once
void aps_fold_local_cuda(const float* const* srcs, int W, float* out, const Chunk* chunks,
                         int n_chunks, int T, int T_cap, int e, int m, int rounding, bool use_aps, bool kahan,
                         float world_mult, unsigned* workspace, int* shifts_out, int max_ctas,
                         cudaStream_t st);

void lars_step_cuda(float* master, float* model, const float* grad, float* mom, const Chunk* chunks,
                    const int* tchunk_off, int n_chunks, int T, const float* lr_dev, float lr,
                    float momentum, float wd, float coeff, float eps, bool use_lars, bool nesterov,
                    const float* wd_mult, float* partials, unsigned* barrier_ws, int max_ctas,
                    cudaStream_t st);

// Cross-GPU fused all-reduce over peer memory.  See cuda/allreduce_fused_sm100a.cu.
struct CommBuffers {
  int rank, world;
  // per-rank device pointers into the symmetric region (index = rank)
  unsigned char* data[16];     // wire buffers   [2][cap_bytes]
  unsigned char* gather[16];   // reduced-slice buffers [2][cap_bytes]
  float* mailbox[16];          // APS exponents  [2 parities][8 source ranks][max_tensors] (pushed by the sources)
  unsigned* flags[16];         // signal pad     [3][16] epoch flags per phase
  unsigned char* mc_gather;    // multicast view of `gather` (nullptr when no NVLS)
  unsigned char* mc_data;      // multicast view of `data` (in-switch reductions: multimem.ld_reduce)
  size_t cap_bytes;
  int max_tensors;
};
// ALGO_NVLS_SUM: fp32 wire, the whole reduction inside the switch (multimem.ld_reduce.add.f32) - the (8,23) path, whose
//   reference is a plain NCCL all_reduce (dist_util.py:55-59; summation order unspecified there too).
// ALGO_NVLS_FAST: fp8 wire codes summed inside the switch with an fp16 accumulator (multimem.ld_reduce ... acc::f16):
//   explicitly NOT the rank-ordered CPD fold (one rounding at the end instead of one per add) - opt-in only.
enum AllreduceAlgo : int { ALGO_AUTO = 0, ALGO_ONESHOT = 1, ALGO_TWOSHOT = 2, ALGO_NVLS = 3, ALGO_ONESHOT_RAW = 4,
                           ALGO_NVLS_SUM = 5, ALGO_NVLS_FAST = 6 };
void allreduce_fused_cuda(const CommBuffers& cb, float* grad, const Chunk* chunks, int n_chunks,
                          int T, int e, int m, int rounding, bool use_aps, bool kahan, int algo, bool prequant, int wire,
                          unsigned* local_ws /*barrier[2], epoch, status, amax[2][T]*/, int* shifts_out,
                          int max_ctas, unsigned long long timeout_cycles, cudaStream_t st);

// Operand staging for the tensor-core GEMM (cuda/operand_stage_sm100.cu): gather + cast in one kernel.
struct StageDesc {
  const float* src;
  long long base;          // element offset added to every address (flips start from the far end)
  int mode;                // 0 = GATHER (strided, k split in three), 1 = IM2COL
  int rows_fast;           // 1: the source is contiguous along the output ROW index (read r-fast), 0: along k
  int vec4;                // set by the launcher: 4 consecutive fast-dimension elements are one aligned 16-byte load
  // GATHER: address = base + r * s_row + k1 * s1 + k2 * s2 + k3 * s3,  k = (k1 * n2 + k2) * n3 + k3
  long long s_row, s1, s2, s3;
  int n2, n3;
  // IM2COL over x[b, ih, iw, c] with element strides (sb, sh, sw, sc): pixel = (b, oh, ow), tap = (kh, kw, c)
  int taps_are_rows;       // 0: rows = pixels, k = taps;  1: rows = taps, k = pixels (transposed im2col)
  int C, H, W, OH, OW, KH, KW, stride, pad, dil, in_dil;
  long long sb, sh, sw, sc;
};
// operand_mode as in gemm_tc_cuda (0 bf16, 1 fp32/tf32, 2 e4m3, 3 e5m2); dst rows have Kp elements (Kp % 8 == 0)
int operand_stage_cuda(const StageDesc& d, void* dst, long long R, long long K, long long Kp, int operand_mode, cudaStream_t st);
// both operands of a GEMM in ONE launch (n = 1 or 2)
int operand_stage2_cuda(const StageDesc* descs, void* const* dsts, const long long* R, const long long* K, const long long* Kp,
                        int n, int operand_mode, cudaStream_t st);
// out[j] = Q(sum_i x[i*ld + j]) : the bias gradient of Quant_Linear / Quant_Conv
void colsum_quantize_cuda(const float* x, long long rows, int cols, long long ld, float* out, int e, int m, int rounding,
                          cudaStream_t st);

// Training-mode BatchNorm2d fused with the residual add and ReLU that follow it (cuda/bn_act_sm100.cu), channels_last fp32:
// x, res, y, dy, dx, dres are [M = N*H*W, C] row-major, C % 4 == 0.  stats = mean | rstd | scale | bias (4*C floats),
// grads = dgamma | dbeta | c1 | c2 (4*C floats); workspace: bn_workspace_floats(M, C) floats, zero-initialised once.
long long bn_workspace_floats(long long M, int C);
int bn_act_forward_cuda(const float* x, const float* res, float* y, long long M, int C, const float* gamma, const float* beta,
                        float* running_mean, float* running_var, long long* num_batches_tracked, float momentum, float eps,
                        bool relu, float* stats, float* workspace, unsigned* relu_mask, cudaStream_t st);
// relu_mask (optional, M*C/32 words, M*C % 32 == 0): 1 bit per element written by the forward apply; the backward then
// reads it instead of y (either y or relu_mask must be given when relu)
int bn_act_backward_cuda(const float* dy, const float* y, const unsigned* relu_mask, const float* x, float* dx, float* dres, long long M,
                         int C, const float* gamma, const float* stats, bool relu, float* grads, float* workspace, cudaStream_t st);

// tcgen05 / TMEM / TMA GEMM with per-K-slice accumulator re-quantisation.
// operand_mode: 0 = bf16 operands (kind::f16), 1 = tf32 (kind::tf32), 2 = e4m3, 3 = e5m2
// variant: 0 = auto, 1 = 128x128 single-CTA tiles, 2 = 256x256 CTA-pair tiles (cta_group::2 cluster kernel,
// packed 16-bit accumulators; formats (5,10), (8,7), (4,3), (5,2) only)
// lda / ldb: row strides of A / Bt in ELEMENTS (0 = K; staged operands have padded rows, the TMA descriptor still
// declares the true K so the tail is zero-filled in flight).  splits > 1: split-K — split z accumulates its own
// contiguous range of K-slices into workspace[z] (fp32 [splits, M, N]); a second kernel folds the splits IN ORDER,
// C = fold_z Q(C (+) part_z) (+ bias): the ordered low-precision fold of CPD applied across K ranges, which gives
// the skinny weight-gradient GEMMs (reduction over batch x pixels) a grid that fills the machine.
// conv != nullptr: IMPLICIT-GEMM convolution - A is not a matrix but the NHWC activation tensor itself (already in the
// operand dtype), read through an im2col-mode TMA descriptor: M = B*OH*OW output positions, K = KH*KW*C in (kh, kw, c)
// order; lda is ignored.  Needs C % (128 B / element size) == 0.
struct ConvGeom {
  int B, H, W, C;               // activation tensor [B, H, W, C], C contiguous
  int KH, KW, OH, OW;
  int stride, pad, dil;
};
int gemm_tc_cuda(const void* A, const void* B, float* C, int M, int N, int K, int e, int m,
                 bool kahan, int k_slice, int operand_mode, const float* bias, int variant, cudaStream_t st,
                 long long lda = 0, long long ldb = 0, int splits = 1, float* workspace = nullptr, const ConvGeom* conv = nullptr);
// MXFP8 operand staging: fp32 [R,K] -> fp8 codes [R,Kp] + UE8M0 per-32 block scales in the tcgen05 scale-factor atom layout
void mx_quantize_cuda(const float* in, long long ld_in, uint8_t* codes, uint8_t* sf, int R, int K, int Kp, bool e5m2, cudaStream_t st);
// block-scaled GEMM (tcgen05.mma.kind::mxf8f6f4.block_scale) with per-K-slice accumulator requant
int gemm_tc_mx_cuda(const uint8_t* codes_a, const uint8_t* sf_a, const uint8_t* codes_b, const uint8_t* sf_b, float* C, int M, int N,
                    int Kp, int e, int m, bool kahan, int k_slice, bool e5m2, const float* bias, cudaStream_t st);
// weight gradient of a convolution as an implicit GEMM (MN-major operand tiles; bf16 operands): see gemm_tcgen05_sm100a.cu
int gemm_wgrad_cuda(const void* dY, const void* X, float* dW, const ConvGeom& g, int O, int e, int m, bool kahan, int k_slice,
                    cudaStream_t st, int splits = 1, float* workspace = nullptr, int lbo = 0, int sbo = 0);
int gemm_tc2_cuda(const void* A, const void* Bt, float* C, int M, int N, int K, int e, int m, bool kahan, int k_slice,
                  int operand_mode, const float* bias, cudaStream_t st, long long lda = 0, long long ldb = 0,
                  const ConvGeom* conv = nullptr);
#endif

}  // namespace cpd
