// common.cuh -- shared device helpers (reductions, async copies, error plumbing).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/admm_lstm_b200.h"
#include "admm_math.cuh"

namespace admm {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_launch(const char* what);

// Per-kernel CUDA-event timing (admm_kernel_timing in the C ABI; bench.py's per-kernel roofline table).  A KernelScope placed
// before a launch records an event pair around everything launched on `st` during its lifetime when timing is enabled;
// disabled (the default) it costs one relaxed atomic load.
struct KernelScope {
  KernelScope(const char* name, cudaStream_t st);
  ~KernelScope();
  KernelScope(const KernelScope&) = delete;
  KernelScope& operator=(const KernelScope&) = delete;
 private:
  const char* name_;
  cudaStream_t st_;
  cudaEvent_t e0_;
};

#define ADMM_REQUIRE(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      ::admm::set_error(__VA_ARGS__);    \
      return ADMM_EINVAL;                \
    }                                    \
  } while (0)

__host__ __device__ inline Rho make_rho(const admm_hyper& hp) {
  Rho r;
  r.i = hp.rho[0]; r.f = hp.rho[1]; r.g = hp.rho[2]; r.o = hp.rho[3];
  r.c = hp.rho[4]; r.h = hp.rho[5]; r.y = hp.rho[6];
  return r;
}

// Low part of the 3xTF32 split.  The tensor core truncates the raw fp32 operand to tf32 itself (a_hi), so
// a_lo = a - trunc_tf32(a) exactly; a_lo is then rounded to nearest tf32 here, because the MMA would
// otherwise truncate it too and bias every product towards zero.
__device__ __forceinline__ float tf32_round(float v) {
  return __uint_as_float((__float_as_uint(v) + 0x00001000u) & 0xFFFFE000u);
}
__device__ __forceinline__ float tf32_lo(float a) {
  return tf32_round(a - __uint_as_float(__float_as_uint(a) & 0xFFFFE000u));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of NV per-thread fp32 partials -> one fp64 atomicAdd per value per CTA.
// `red` is shared scratch of at least NV * (blockDim/32) floats.
template <int NV>
__device__ __forceinline__ void block_accumulate(const float (&v)[NV], float* red, double* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const float s = warp_sum(v[k]);
    if (lane == 0) red[k * nwarp + warp] = s;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < NV; k += blockDim.x) {
    double s = 0.0;
    for (int w = 0; w < nwarp; ++w) s += (double)red[k * nwarp + w];
    atomicAdd(out + k, s);
  }
  __syncthreads();
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Streaming (read-once / write-once) 128-bit accesses for the state tensors: keep them out of L1.
__device__ __forceinline__ float4 ld_stream(const float* p) {
  float4 v;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];\n"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream(float* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};\n"
               ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Scalar streaming accesses as volatile asm: ptxas keeps volatile asm statements in program order, which pins the
// epilogue schedule to [all loads of a batch][arithmetic][all stores] (the intrinsics let it sink every load next to
// its first use, serialising the memory latency).
__device__ __forceinline__ float ldg_stream(const float* p) {
  float v;
  asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream(float* p, float v) {
  asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void stg_keep(float* p, float v) {
  asm volatile("st.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

}  // namespace admm
