// grad_from_z.cu -- gradient residual of the x-phase weight update from STORED pre-activations.
//
// The sweep of iteration s-1 (and the forward initialisation) runs z_t = x_t W + h_{t-1} U for every t with exactly
// the weights and states that iteration s starts from, so its GEMM result is kept (zstore) and the x-phase gradient
// pass of iteration s (admm.py:302-312) needs no GEMM: per element
//      u = act(z) - lambda/rho - gate,   R = u * act'(z),   f(w) += u^2   (admm.py:316-325 at beta = w)
// R^T (and its tf32 low part) goes to the scratch layout the A^T R reduction GEMM reads.  Pure streaming kernel:
// 12 B read, 8 B written per (gate, unit, sample, timestep); float4 accesses, one fp64 atomic per CTA.
#include "common.cuh"
#include "gate_gemm.h"

namespace admm {
namespace {

constexpr int NT = 256;

// same operand scaling as the tensor-core epilogues (gate_gemm_tc.cu cap_exp / split_f16)
__device__ __forceinline__ int g_cap_exp(unsigned max_bits) {
  const float m = __uint_as_float(max_bits);
  if (!(m > 0.f) || !isfinite(m)) return 0;
  int e;
  frexpf(m, &e);
  return 13 - e;
}
__device__ __forceinline__ void g_split(float vs, __half* hi, __half* lo) {
  const float c = fminf(fmaxf(vs, -65504.0f), 65504.0f);
  const __half h = __float2half_rn(c);
  *hi = h;
  *lo = __float2half_rn(c - __half2float(h));
}

__global__ void __launch_bounds__(NT) grad_from_z_kernel(const GradFromZArgs p, int n_nb, int64_t items_per_gate) {
  __shared__ float red[NT / 32];
  if (p.skip_if && *p.skip_if != 0) return;
  const int g = blockIdx.y;
  const float rho = p.rho[g];
  float fsum = 0.f, bmax = 0.f;
  const float inv_rho = 1.0f / rho;
  const float r_scale = p.r16_hi ? ldexpf(1.0f, g_cap_exp(*p.r_bound)) : 1.0f;
  for (int64_t item = blockIdx.x; item < items_per_gate; item += gridDim.x) {
    const int nb = (int)(item % n_nb);
    const int64_t row = item / n_nb;                 // (unit j, timestep tl)
    const int tl = (int)(row % p.tc), j = (int)(row / p.tc);
    const int64_t n = ((int64_t)nb * NT + threadIdx.x) * 4;
    if (n >= p.ldn) continue;
    const int64_t zo = (((int64_t)g * p.H + j) * p.zT + p.zt0 + tl) * p.ldn + n;
    const int64_t so = (int64_t)tl * p.s_tstride + (int64_t)j * p.ldn + n;
    const int64_t ro = (((int64_t)g * p.H + j) * p.tc + tl) * p.ldn + n;
    const float4 z4 = ld_stream(p.zstore + zo), lam4 = ld_stream(p.dual[g] + so), gv4 = ld_stream(p.gate[g] + so);
    const float z[4] = {z4.x, z4.y, z4.z, z4.w}, lam[4] = {lam4.x, lam4.y, lam4.z, lam4.w};
    const float gv[4] = {gv4.x, gv4.y, gv4.z, gv4.w};
    float r[4], rl[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float u;
      const float rr = grad_point<FastMath>(z[e], lam[e], gv[e], rho, g == 2, &u);
      const bool ok = n + e < p.n;
      r[e] = ok ? rr : 0.f;
      rl[e] = tf32_lo(r[e]);
      if (ok) fsum = fmaf(u, u, fsum);
      bmax = fmaxf(bmax, 1.0f + fabsf(lam[e]) * inv_rho + fabsf(gv[e]));     // >= |u| >= |R| for any z
    }
    if (p.r16_hi) {
      __align__(8) __half hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) g_split(r[e] * r_scale, &hi[e], &lo[e]);
      *reinterpret_cast<uint2*>(p.r16_hi + ro) = *reinterpret_cast<const uint2*>(hi);      // re-read at once by atr: keep in L2
      *reinterpret_cast<uint2*>(p.r16_lo + ro) = *reinterpret_cast<const uint2*>(lo);
    } else {
      *reinterpret_cast<float4*>(p.r + ro) = make_float4(r[0], r[1], r[2], r[3]);
      *reinterpret_cast<float4*>(p.r_lo + ro) = make_float4(rl[0], rl[1], rl[2], rl[3]);
    }
  }
  if (p.bound_track) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bmax = fmaxf(bmax, __shfl_xor_sync(0xffffffffu, bmax, o));
    if ((threadIdx.x & 31) == 0) atomicMax(p.bound_track, __float_as_uint(bmax));
  }
  const float s = warp_sum(fsum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double acc = 0.0;
    for (int w = 0; w < NT / 32; ++w) acc += (double)red[w];
    atomicAdd(p.fw_acc + g, acc);
  }
}

}  // namespace

int grad_from_z(const GradFromZArgs& a, cudaStream_t st) {
  const int n_nb = (int)((a.ldn / 4 + NT - 1) / NT);
  const int64_t items = (int64_t)a.H * a.tc * n_nb;
  const unsigned gx = (unsigned)(items < 148 * 2 ? items : 148 * 2);
  KernelScope ks_("grad_from_z_kernel", st);
  grad_from_z_kernel<<<dim3(gx, 4), NT, 0, st>>>(a, n_nb, items);
  count_launch();
  return check_launch("grad_from_z");
}

}  // namespace admm
