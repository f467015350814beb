// device_utils.cuh — small device-side building blocks shared by the sm_100a
// kernels: 128-bit streaming loads/stores, system-scope release/acquire flag
// access for cross-GPU signalling, a reusable (graph-replay safe) grid barrier,
// block reductions.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace cpd {

// ---- vector memory access ----------------------------------------------------
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld_stream_f4(const void* p) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
// plain (cacheable) 128-bit load — used for data re-read from L2 inside one kernel
__device__ __forceinline__ float4 ld_f4(const void* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st_f4(void* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st_stream_f4(void* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_u4(void* p, uint4 v) { *reinterpret_cast<uint4*>(p) = v; }
// peer / system-visible data: volatile-relaxed at system scope so the value is
// fetched from the owner (never a stale L1 line of a previous epoch)
__device__ __forceinline__ uint4 ld_sys_u4(const void* p) {
  uint4 r;
  asm volatile("ld.relaxed.sys.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ uint2 ld_sys_u2(const void* p) {
  uint2 r;
  asm volatile("ld.relaxed.sys.global.L1::no_allocate.v2.u32 {%0,%1}, [%2];"
               : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ float ld_sys_f32(const void* p) {
  float r;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(r) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void st_sys_u4(void* p, uint4 v) {
  asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_sys_u2(void* p, uint2 v) {
  asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
}
// NVLS: one store, replicated by the switch into every GPU bound to the multicast object
__device__ __forceinline__ void multimem_st_u4(void* mc, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(mc), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)),
                  "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w)) : "memory");
}
__device__ __forceinline__ void multimem_st_u2(void* mc, uint2 v) {
  asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1,%2};"
               :: "l"(mc), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)) : "memory");
}
// in-switch fp32 sum of the same address on every GPU (plain-fp32 all-reduce mode)
__device__ __forceinline__ float4 multimem_ld_reduce_add_f4(const void* mc) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc) : "memory");
  return r;
}

__device__ __forceinline__ void multimem_st_f4(void* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// in-switch sum of the 16 fp8 codes at the same address on every GPU, accumulated in fp16 inside the switch and rounded
// back to fp8 codes once (LDGMC.E.HPADD.E4M3x16 / E5M2x16): ONE rounding per element instead of the CPD fold's one per
// add - the explicitly non-bit-exact "nvls_fast" mode
template <bool E5M2>
__device__ __forceinline__ uint4 multimem_ld_reduce_add_fp8x16(const void* mc) {
  uint4 r;
  if (E5M2)
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f16.v4.e5m2x4 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(mc) : "memory");
  else
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f16.v4.e4m3x4 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(mc) : "memory");
  return r;
}

// ---- flags -------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// ---- grid barrier --------------------------------------------------------------
// bar[0] = arrival count (returns to 0 after every barrier), bar[1] = generation
// (monotonic, never reset) — safe to reuse across launches and CUDA-graph replays
// without host-side resets.  All CTAs of the grid must be co-resident.
__device__ __forceinline__ void grid_barrier(unsigned* bar) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned gen = ld_acquire_gpu(&bar[1]);
    __threadfence();
    const unsigned arrived = atomicAdd(&bar[0], 1u) + 1u;
    if (arrived == gridDim.x) {
      bar[0] = 0u;
      __threadfence();
      red_release_gpu_add(&bar[1], 1u);
    } else {
      while (ld_acquire_gpu(&bar[1]) == gen) { __nanosleep(20); }
    }
    __threadfence();
  }
  __syncthreads();
}

// ---- block reductions ----------------------------------------------------------
__device__ __forceinline__ unsigned warp_max_u32(unsigned v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum_f32(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// smem must hold >= 32 entries; result valid in every thread
__device__ __forceinline__ unsigned block_max_u32(unsigned v, unsigned* smem) {
  v = warp_max_u32(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  unsigned r = lane < nw ? smem[lane] : 0u;
  return warp_max_u32(r);
}
__device__ __forceinline__ float block_sum_f32(float v, float* smem) {
  v = warp_sum_f32(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  float r = lane < nw ? smem[lane] : 0.f;
  return warp_sum_f32(r);
}

}  // namespace cpd
