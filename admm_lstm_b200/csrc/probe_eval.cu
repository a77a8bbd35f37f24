// probe_eval.cu -- the backtracking probes of the weight update (admm.py:316-325, 331-336) for a whole vector
// of candidate thetas in ONE pass over the data:
//
//   f_k = sum_{t,n,j} ( act(Z0 + Q * 2^-(k0+k)) - lambda/rho - gate )^2 ,   k = 0..ncand-1,   and f(w) at Q = 0
//
// Z0 = A w + B w' and Q = A_src G come from the gate GEMM (PROBE mode).  beta_k = w + G/theta_k, so
// A beta_k + B w' = Z0 + Q/theta_k: the reference re-runs the full GEMM for every probe (and f(w) again inside
// every est()), this kernel re-uses one GEMM for all of them.  Pure elementwise work at full occupancy:
// float4 loads, per-thread fp32 partials, one fp64 atomic per (block, candidate).
//
// Activations: the kernel evaluates (ncand+1) * 4H activations per sample-timestep, so they are hand-rolled to
// two MUFU ops each (ex2.approx + rcp.approx with one Newton step) plus a degree-5 odd polynomial for
// |x| < 0.6 in tanh, where the exponential form loses relative accuracy.  Max error ~2 ulp, the same class as
// expf()/tanhf() (a plain ex2/rcp tanh without the polynomial was measured to flip backtracking decisions whenever
// a constraint residual sits at the fp32 rounding level, moving the trajectory by ~1e-3).  The sums only feed
// the comparison f(beta) > est; f(w) comes from the same functions; state is never written from them.
#include "common.cuh"
#include "gate_gemm.h"

namespace admm {
namespace {

constexpr int NT = 256;
constexpr int JB = 8;          // hidden units per block
constexpr int NCS = ADMM_FK_SLOTS;

// Residual of one element for one candidate.  The instruction stream is what bounds this kernel (ncu: issue slots
// 69 % busy, XU pipe 56 %, at 22 instructions per activation before this was flattened), so the sigmoid path is
// spelled out: FFMA, FMNMX, FMUL, MUFU.EX2, FADD, MUFU.RCP, 2 FFMA (Newton), 2 FADD, FFMA = 11 issue slots.
// 1/d for d in [1, 2^116] on the FMA pipe only: bit-trick seed (12 % off) + three Newton steps (1.4e-2, 2e-4, 4e-8).
// Every second candidate uses it instead of MUFU.RCP: the kernel is bound by the XU pipe (2 MUFU = 16 XU cycles per
// warp-activation against 11 issue slots), so trading one MUFU for five FMA-pipe instructions on half of the
// candidates balances the two pipes (12 XU cycles against 13.5 issue slots on average).
__device__ __forceinline__ float rcp_fma(float d) {
  float r = __uint_as_float(0x7EF311C7u - __float_as_uint(d));
  r = r * fmaf(-d, r, 2.0f);
  r = r * fmaf(-d, r, 2.0f);
  return fmaf(r, fmaf(-d, r, 1.0f), r);
}
template <bool IS_G>
__device__ __forceinline__ float probe_term(float z, float lr, float gv, bool fma_rcp) {   // fma_rcp folds after unrolling
  float a;
  if (IS_G) {
    a = FastMath::tanh(z);
  } else if (fma_rcp) {
    a = rcp_fma(1.0f + FastMath::ex2(-1.4426950408889634f * fmaxf(z, -80.0f)));
  } else {
    a = FastMath::sigmoid(z);
  }
  return (a - lr) - gv;                 // the reference's association (admm.py:319-323)
}

// NSLOT accumulators: candidates [0, NC) at theta = 2^-(kbase + k) and, if WITH_FW, slot NC at Q = 0 (= f(w)).
// FULL = all four samples of the thread are real (no ghost-row masking).
template <int NC, bool WITH_FW, bool IS_G, bool FULL>
__device__ __forceinline__ void accumulate(const float4& z4, const float4& q4, const float4& lam4, const float4& gv4,
                                           float rho, float inv_rho, bool rho_pow2, const float (&inv_theta)[NC],
                                           const float (&mask)[4], float (&acc)[NC + (WITH_FW ? 1 : 0)]) {
  const float z[4] = {z4.x, z4.y, z4.z, z4.w}, q[4] = {q4.x, q4.y, q4.z, q4.w};
  const float lam[4] = {lam4.x, lam4.y, lam4.z, lam4.w}, gv[4] = {gv4.x, gv4.y, gv4.z, gv4.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    // lambda / rho: exact through the reciprocal when rho is a power of two (the shipped gate rho are 1 except HAR 1.5, PTB 0.8, UCF101 0.1)
    const float lr = rho_pow2 ? lam[e] * inv_rho : __fdiv_rn(lam[e], rho);
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      float u = probe_term<IS_G>(fmaf(q[e], inv_theta[k], z[e]), lr, gv[e], (k & 1) != 0);
      if (!FULL) u *= mask[e];
      acc[k] = fmaf(u, u, acc[k]);
    }
    if (WITH_FW) {
      float u = probe_term<IS_G>(z[e], lr, gv[e], false);
      if (!FULL) u *= mask[e];
      acc[NC] = fmaf(u, u, acc[NC]);
    }
  }
}

// Work item = (gate g, timestep tl, block of JB units, block of 4*NT samples).  The grid is a fixed number of
// CTAs that stride over the items: a launch whose gates are all decided (the speculative later passes of the
// backtracking) then costs a few hundred CTAs that exit at once instead of one CTA per item.
template <int NC, bool WITH_FW>
__global__ void __launch_bounds__(NT) probe_eval_kernel(const ProbeEvalArgs p, int n_jb, int n_nb, int64_t n_items) {
  constexpr int NSLOT = NC + (WITH_FW ? 1 : 0);
  __shared__ float red[NSLOT * (NT / 32)];
  if (p.done[0] && p.done[1] && p.done[2] && p.done[3]) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int nb = (int)(item % n_nb);
    const int jb = (int)((item / n_nb) % n_jb) * p.jmod + p.jrem;
    const int gt = (int)(item / ((int64_t)n_nb * n_jb));
    const int g = gt & 3, tl = gt >> 2;
    if (p.done[g]) continue;
    const int kbase = p.kbase[g], nc = p.nc[g];
    if (nc <= 0) continue;
    float inv_theta[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) inv_theta[k] = (k < nc) ? __int_as_float((127 - (kbase + k)) << 23) : 0.f;   // 2^-(kbase+k), kbase+k < 64
    const int64_t n = ((int64_t)nb * NT + threadIdx.x) * 4;
    const int j0 = jb * JB;
    float acc[NSLOT];
#pragma unroll
    for (int k = 0; k < NSLOT; ++k) acc[k] = 0.f;
    const float rho = p.rho[g];
    const float inv_rho = 1.0f / rho;
    const bool rho_pow2 = (__float_as_uint(rho) & 0x007FFFFFu) == 0u;
    const bool block_full = ((int64_t)(nb + 1) * NT * 4 <= p.n);       // uniform over the CTA
    if (n < p.ldn) {
      const float mask[4] = {n + 0 < p.n ? 1.f : 0.f, n + 1 < p.n ? 1.f : 0.f, n + 2 < p.n ? 1.f : 0.f,
                             n + 3 < p.n ? 1.f : 0.f};
      const float* gate = p.gate[g] + (int64_t)tl * p.s_tstride + n;
      const float* dual = p.dual[g] + (int64_t)tl * p.s_tstride + n;
      const int jend = (p.H - j0 < JB) ? p.H - j0 : JB;
#pragma unroll 2
      for (int jj = 0; jj < jend; ++jj) {
        const int j = j0 + jj;
        const int64_t so = (((int64_t)g * p.H + j) * p.tc + tl) * p.ldn + n;
        const int64_t zo = (((int64_t)g * p.H + j) * p.z_T + p.z_t0 + tl) * p.ldn + n;
        const float4 z4 = ld_stream(p.z0 + zo), q4 = ld_stream(p.q + so);
        const float4 lam4 = ld_stream(dual + (int64_t)j * p.ldn), gv4 = ld_stream(gate + (int64_t)j * p.ldn);
        if (g == 2) {
          if (block_full) accumulate<NC, WITH_FW, true, true>(z4, q4, lam4, gv4, rho, inv_rho, rho_pow2, inv_theta, mask, acc);
          else accumulate<NC, WITH_FW, true, false>(z4, q4, lam4, gv4, rho, inv_rho, rho_pow2, inv_theta, mask, acc);
        } else {
          if (block_full) accumulate<NC, WITH_FW, false, true>(z4, q4, lam4, gv4, rho, inv_rho, rho_pow2, inv_theta, mask, acc);
          else accumulate<NC, WITH_FW, false, false>(z4, q4, lam4, gv4, rho, inv_rho, rho_pow2, inv_theta, mask, acc);
        }
      }
    }
    // slots in [nc, NC) were evaluated at Q*0 and are not published
#pragma unroll
    for (int k = 0; k < NSLOT; ++k) {
      const float s = warp_sum(acc[k]);
      if (lane == 0) red[k * (NT / 32) + warp] = s;
    }
    __syncthreads();
    if (threadIdx.x < NSLOT) {
      const int k = threadIdx.x;
      if (k < nc || k == NC) {
        double s = 0.0;
        for (int w = 0; w < NT / 32; ++w) s += (double)red[k * (NT / 32) + w];
        const int slot = (k == NC) ? ADMM_MAX_CAND : p.slot0[g] + k;
        atomicAdd(p.fk_acc + g * NCS + slot, s);
      }
    }
    __syncthreads();
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Moment pass (admm_probe_plan::moments).  Along the probe ray the perturbation of the pre-activation is delta = Q 2^-k;
// at and above the exit of the backtracking loop it is tiny (measured 1e-4 .. 1e-7 on the benchmark workloads), so
//     u_k = u + a1 d + ... + a6 d^6 + O(d^7),   a_j = act^(j)(z) / j!,   u = act(z) - lambda/rho - gate
//     u_k^2 - u^2 = c1 d + c2 d^2 + ... + c6 d^6 + O(d^7)
// and the sums B_j = sum c_j Q^j give f for EVERY candidate at once -- one activation per element instead of one per
// element and candidate.  Q is normalised by 2^-k0 (t = Q 2^-k0, |t| <= 2^-4 where the expansion is used) so that the
// fp32 partial sums stay in range; the block totals are rescaled by 2^(j k0) in fp64.
__global__ void __launch_bounds__(NT) probe_moments_kernel(const ProbeEvalArgs p, float* qmax, int n_jb, int n_nb,
                                                           int64_t n_items) {
  __shared__ float red[8 * (NT / 32)];
  if (p.done[0] && p.done[1] && p.done[2] && p.done[3]) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int nb = (int)(item % n_nb);
    const int jb = (int)((item / n_nb) % n_jb);
    const int gt = (int)(item / ((int64_t)n_nb * n_jb));
    const int g = gt & 3, tl = gt >> 2;
    if (p.done[g]) continue;
    const int kn = p.kbase[g];
    const float sn = __int_as_float((127 - kn) << 23);                  // 2^-k0
    const int64_t n = ((int64_t)nb * NT + threadIdx.x) * 4;
    const int j0 = jb * JB;
    float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float tmax = 0.f;
    const float rho = p.rho[g];
    const float inv_rho = 1.0f / rho;
    const bool rho_pow2 = (__float_as_uint(rho) & 0x007FFFFFu) == 0u;
    if (n < p.ldn) {
      const float* gate = p.gate[g] + (int64_t)tl * p.s_tstride + n;
      const float* dual = p.dual[g] + (int64_t)tl * p.s_tstride + n;
      const int jend = (p.H - j0 < JB) ? p.H - j0 : JB;
#pragma unroll 2
      for (int jj = 0; jj < jend; ++jj) {
        const int j = j0 + jj;
        const int64_t so = (((int64_t)g * p.H + j) * p.tc + tl) * p.ldn + n;
        const int64_t zo = (((int64_t)g * p.H + j) * p.z_T + p.z_t0 + tl) * p.ldn + n;
        const float4 z4 = ld_stream(p.z0 + zo), q4 = ld_stream(p.q + so);
        const float4 lam4 = ld_stream(dual + (int64_t)j * p.ldn), gv4 = ld_stream(gate + (int64_t)j * p.ldn);
        const float z[4] = {z4.x, z4.y, z4.z, z4.w}, q[4] = {q4.x, q4.y, q4.z, q4.w};
        const float lam[4] = {lam4.x, lam4.y, lam4.z, lam4.w}, gv[4] = {gv4.x, gv4.y, gv4.z, gv4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (n + e >= p.n) continue;
          const float lr = rho_pow2 ? lam[e] * inv_rho : __fdiv_rn(lam[e], rho);
          const float s = (g == 2) ? FastMath::tanh(z[e]) : FastMath::sigmoid(z[e]);
          const float u = (s - lr) - gv[e];
          float c[6];
          moment_terms(g == 2, s, u, c);
          const float t = q[e] * sn, t2 = t * t, t3 = t2 * t;
          acc[0] = fmaf(u, u, acc[0]);
          acc[1] = fmaf(c[0], t, acc[1]);
          acc[2] = fmaf(c[1], t2, acc[2]);
          acc[3] = fmaf(c[2], t3, acc[3]);
          acc[4] = fmaf(c[3] * t2, t2, acc[4]);
          acc[5] = fmaf(c[4] * t2, t3, acc[5]);
          acc[6] = fmaf(c[5] * t3, t3, acc[6]);
          tmax = fmaxf(tmax, fabsf(q[e]));
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const float sm = warp_sum(acc[k]);
      if (lane == 0) red[k * (NT / 32) + warp] = sm;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
    if (lane == 0) red[7 * (NT / 32) + warp] = tmax;
    __syncthreads();
    if (threadIdx.x < 7) {
      const int k = threadIdx.x;
      double sm = 0.0;
      for (int w = 0; w < NT / 32; ++w) sm += (double)red[k * (NT / 32) + w];
      atomicAdd(p.fk_acc + g * NCS + ADMM_FK_MOMENTS + k, ldexp(sm, k * kn));      // B_k in units of Q^k, not t^k
    } else if (threadIdx.x == 7) {
      float m = 0.f;
      for (int w = 0; w < NT / 32; ++w) m = fmaxf(m, red[7 * (NT / 32) + w]);
      atomicMax(reinterpret_cast<unsigned int*>(qmax + g), __float_as_uint(m));
    }
    __syncthreads();
  }
}

template <int NC>
void launch_probe_eval(const ProbeEvalArgs& a, unsigned grid, int n_jb, int n_nb, int64_t n_items, cudaStream_t st) {
  if (a.publish_fw) probe_eval_kernel<NC, true><<<grid, NT, 0, st>>>(a, n_jb, n_nb, n_items);
  else probe_eval_kernel<NC, false><<<grid, NT, 0, st>>>(a, n_jb, n_nb, n_items);
}

}  // namespace

int probe_eval(const ProbeEvalArgs& a, cudaStream_t st) {
  const int n_nb = (int)((a.ldn / 4 + NT - 1) / NT);
  const int n_jb_all = (a.H + JB - 1) / JB;
  const int n_jb = (n_jb_all - a.jrem + a.jmod - 1) / a.jmod;
  const int nc = max(max(a.nc[0], a.nc[1]), max(a.nc[2], a.nc[3]));
  if (nc <= 0 || n_jb <= 0) return ADMM_OK;
  const int64_t n_items = (int64_t)n_nb * n_jb * 4 * a.tc;
  const unsigned grid = (unsigned)(n_items < 148 * 8 ? n_items : 148 * 8);
  KernelScope ks_("probe_eval_kernel", st);
  if (nc <= 4) launch_probe_eval<4>(a, grid, n_jb, n_nb, n_items, st);
  else if (nc <= 8) launch_probe_eval<8>(a, grid, n_jb, n_nb, n_items, st);
  else if (nc <= 12) launch_probe_eval<12>(a, grid, n_jb, n_nb, n_items, st);
  else if (nc <= 16) launch_probe_eval<16>(a, grid, n_jb, n_nb, n_items, st);
  else if (nc <= 24) launch_probe_eval<24>(a, grid, n_jb, n_nb, n_items, st);
  else launch_probe_eval<ADMM_MAX_CAND>(a, grid, n_jb, n_nb, n_items, st);
  count_launch();
  return check_launch("probe_eval");
}

int probe_moments(const ProbeEvalArgs& a, float* qmax, cudaStream_t st) {
  const int n_nb = (int)((a.ldn / 4 + NT - 1) / NT);
  const int n_jb = (a.H + JB - 1) / JB;
  const int64_t n_items = (int64_t)n_nb * n_jb * 4 * a.tc;
  const unsigned grid = (unsigned)(n_items < 148 * 8 ? n_items : 148 * 8);
  KernelScope ks_("probe_moments_kernel", st);
  probe_moments_kernel<<<grid, NT, 0, st>>>(a, qmax, n_jb, n_nb, n_items);
  count_launch();
  return check_launch("probe_moments");
}

}  // namespace admm
