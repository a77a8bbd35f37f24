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

// One (gate, unit, timestep) row = ldn contiguous samples of z / lambda / gate.  A CTA walks rows (grid stride) and, inside a
// row, blocks of NT float4; the row decomposition is the only integer division (32-bit, once per row, CTA-uniform) and two
// blocks are in flight per thread.  IS_G / F16 are compile-time: one activation per element (a runtime `g == 2 ? tanh :
// sigmoid` evaluates both: 67 instructions per element, issue-bound under the power cap; now ~30, HBM-bound).
struct RowPtrs {
  const float* z;
  const float* lam;
  const float* gate;
  int64_t ro;
};

template <bool IS_G, bool F16>
__device__ __forceinline__ void gfz_block(const GradFromZArgs& p, const RowPtrs& rp, int64_t n, const float4& z4, const float4& lam4,
                                          const float4& gv4, float rho, float inv_rho, float r_scale, float& fsum, float& bmax) {
  const float z[4] = {z4.x, z4.y, z4.z, z4.w}, lam[4] = {lam4.x, lam4.y, lam4.z, lam4.w};
  const float gv[4] = {gv4.x, gv4.y, gv4.z, gv4.w};
  float r[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float u;
    const float rr = grad_point<FastMath>(z[e], lam[e], gv[e], rho, IS_G, &u);
    const bool ok = n + e < p.n;
    r[e] = ok ? rr : 0.f;
    if (ok) fsum = fmaf(u, u, fsum);
    bmax = fmaxf(bmax, 1.0f + fabsf(lam[e]) * inv_rho + fabsf(gv[e]));     // >= |u| >= |R| for any z
  }
  const int64_t ro = rp.ro + n;
  if (F16) {
    __align__(8) __half hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) g_split(r[e] * r_scale, &hi[e], &lo[e]);
    *reinterpret_cast<uint2*>(p.r16_hi + ro) = *reinterpret_cast<const uint2*>(hi);      // re-read at once by atr: keep in L2
    *reinterpret_cast<uint2*>(p.r16_lo + ro) = *reinterpret_cast<const uint2*>(lo);
  } else {
    *reinterpret_cast<float4*>(p.r + ro) = make_float4(r[0], r[1], r[2], r[3]);
    *reinterpret_cast<float4*>(p.r_lo + ro) = make_float4(tf32_lo(r[0]), tf32_lo(r[1]), tf32_lo(r[2]), tf32_lo(r[3]));
  }
}

template <bool IS_G, bool F16>
__device__ __forceinline__ void gfz_rows(const GradFromZArgs& p, int g, int rows, float& fsum, float& bmax) {
  const float rho = p.rho[g];
  const float inv_rho = 1.0f / rho;
  const float r_scale = F16 ? ldexpf(1.0f, g_cap_exp(*p.r_bound)) : 1.0f;
  const int64_t step = (int64_t)NT * 4;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int j = row / p.tc, tl = row - j * p.tc;       // (unit j, timestep tl), tl fastest as in the scratch layout
    RowPtrs rp;
    const int64_t so = (int64_t)tl * p.s_tstride + (int64_t)j * p.ldn;
    rp.z = p.zstore + (((int64_t)g * p.H + j) * p.zT + p.zt0 + tl) * p.ldn;
    rp.lam = p.dual[g] + so;
    rp.gate = p.gate[g] + so;
    rp.ro = (((int64_t)g * p.H + j) * p.tc + tl) * p.ldn;
    int64_t n = (int64_t)threadIdx.x * 4;
    for (; n + step < p.ldn; n += 2 * step) {            // two blocks per thread in flight: all six loads before any use
      const int64_t n2 = n + step;
      const float4 za = ld_stream(rp.z + n), la = ld_stream(rp.lam + n), ga = ld_stream(rp.gate + n);
      const float4 zb = ld_stream(rp.z + n2), lb = ld_stream(rp.lam + n2), gb = ld_stream(rp.gate + n2);
      gfz_block<IS_G, F16>(p, rp, n, za, la, ga, rho, inv_rho, r_scale, fsum, bmax);
      gfz_block<IS_G, F16>(p, rp, n2, zb, lb, gb, rho, inv_rho, r_scale, fsum, bmax);
    }
    if (n < p.ldn) {
      const float4 za = ld_stream(rp.z + n), la = ld_stream(rp.lam + n), ga = ld_stream(rp.gate + n);
      gfz_block<IS_G, F16>(p, rp, n, za, la, ga, rho, inv_rho, r_scale, fsum, bmax);
    }
  }
}

__global__ void __launch_bounds__(NT) grad_from_z_kernel(const GradFromZArgs p, int rows) {
  __shared__ float red[NT / 32];
  if (p.skip_if && *p.skip_if != 0) return;
  const int g = blockIdx.y;
  float fsum = 0.f, bmax = 0.f;
  if (p.r16_hi) {
    if (g == 2) gfz_rows<true, true>(p, g, rows, fsum, bmax);
    else gfz_rows<false, true>(p, g, rows, fsum, bmax);
  } else {
    if (g == 2) gfz_rows<true, false>(p, g, rows, fsum, bmax);
    else gfz_rows<false, false>(p, g, rows, fsum, bmax);
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
  // every CTA resident at once (no partial second wave): grid = SMs x CTAs per SM, the four gates in blockIdx.y
  static int ctas = 0;
  if (ctas == 0) {
    int dev = 0, sms = 148, occ = 4;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, grad_from_z_kernel, NT, 0) != cudaSuccess || occ < 1) occ = 4;
    ctas = sms * occ;
  }
  const int rows = a.H * a.tc;
  int gx = ctas / 4;
  if (gx > rows) gx = rows;
  if (gx < 1) gx = 1;
  KernelScope ks_("grad_from_z_kernel", st);
  grad_from_z_kernel<<<dim3((unsigned)gx, 4), NT, 0, st>>>(a, rows);
  count_launch();
  return check_launch("grad_from_z");
}

}  // namespace admm
