// admm_l.cu -- ADMM-LSTM-L: kernels and C ABI.
//
// Reference: /root/reference/comparison_experiment/admm_l/admm_lstm.py (update functions), driven by
// admm_l/main.py:139-191.  Differences in HOW, not WHAT:
//  * weight phase: the residual -z_t + x_t W + h_{t-1} U - lambda_t/rho is linear in W and U (admm_lstm.py:107-163), so
//    the eight updates need only Gram / right-hand-side sums over samples x timesteps.  They come from ONE packing pass
//    (V_g = z_g + lambda_g/rho and h_{t-1}, five row groups) and two runs of the A^T R reduction GEMM of the main path
//    (tcgen05 on fp16 pairs, fp64 accumulation across chunks); the reference re-runs 2T GEMMs per backtracking probe.
//  * sweep: one gate GEMM per timestep (the reference recomputes x W + h U eight times per timestep) and three
//    elementwise kernels separated by the algorithm's own global reductions (torch.max in update_z / update_zg /
//    update_c and fro_qua(o), admm_lstm.py:168,179,225,230), which are also where sample shards must exchange scalars.
// Ghost rows (n >= N) are never touched and stay zero, so no sum needs masking.
#include <string.h>

#include "common.cuh"
#include "gate_gemm.h"
#include "small_kernels.h"
#include "tc_path.h"

namespace admm {
namespace {

constexpr int NT = 256;

__device__ __forceinline__ float l_sig(float x) { return 1.0f / (1.0f + expf(-x)); }   // torch.sigmoid
__device__ __forceinline__ float appro_sig(float m) { return 0.5f * (1.0f + m) + 0.125f; }   // admm_lstm.py:169
__device__ __forceinline__ float appro_tanh(float m) { return 2.0f * (1.0f + m) + 2.0f; }    // admm_lstm.py:180,226

// Slabs of one slot s (all [H][ldn]) plus the scalars.
struct LSlot {
  int64_t n, ldn;
  int32_t H;
  float* gate[6];        // i f g o c h at slot s
  const float* c_prev;   // c at slot s-1
  float* z[4];
  float* lam_s[4];
  float* lam_p[4];
  float* lam9;
  float* lam10;
  const float* P;        // [4][H][ldn] = x W + h_{s-1} U
  __half* h16_hi;        // fp16 pair of h 2^11 at slot s (the gate GEMM's A operand), or nullptr
  __half* h16_lo;
  unsigned* h_ovf;       // sticky flag: |h| >= 32 did not fit the fp16 pair (tc_h_overflow)
  float* next_max;       // [4]: max |gate_g - lambda_p,g/rho_p| of the values this iteration leaves at slot s, i.e. the
                         // torch.max of update_z / update_zg (admm_lstm.py:168,179) of the NEXT iteration -- measured when
                         // the values are written instead of by a separate pass over the state
  unsigned* bound_track; // tensor-core path: running max of |z_g + lambda_s,g/rho_s| and |h| (bit pattern): the bound
                         // that scales the fp16 operand of the next iteration's Gram / right-hand-side pass
  admm_l_hyper hp;
};

__device__ __forceinline__ void l_store_h_side(const LSlot& p, int64_t idx, float h) {
  if (!p.h16_hi) return;
  if (!(fabsf(h * 2048.0f) <= 65504.0f)) *p.h_ovf = 1u;
  const float c = fminf(fmaxf(h * 2048.0f, -65504.0f), 65504.0f);       // 2^11 = SCALE_H of gate_gemm_tc.cu
  const __half hh = __float2half_rn(c);
  p.h16_hi[idx] = hh;
  p.h16_lo[idx] = __float2half_rn(c - __half2float(hh));
}

__device__ __forceinline__ void track_bound(unsigned* slot, float b) {
  if (!slot) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
  if ((threadIdx.x & 31) == 0) atomicMax(slot, __float_as_uint(b));
}

__device__ __forceinline__ void track_max(float* slot, float b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned int*>(slot), __float_as_uint(b));
}

__device__ __forceinline__ float block_max(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float m = 0.f;
  for (int w = 0; w < NT / 32; ++w) m = fmaxf(m, red[w]);
  return m;
}
__device__ __forceinline__ void atomic_max_nonneg(float* addr, float v) {
  atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));     // bit order == value order for v >= 0
}

// ---------------------------------------------------------------------------------------------------- forward
// main.py:89-100 from P: z, gates, c, h
__global__ void __launch_bounds__(NT) l_forward_kernel(const LSlot p) {
  const int64_t total = (int64_t)p.H * p.ldn;
  float bmax = 0.f;
  float gmax[4] = {0.f, 0.f, 0.f, 0.f};
  for (int j = blockIdx.y; j < p.H; j += gridDim.y)
  for (int64_t n = (int64_t)blockIdx.x * NT + threadIdx.x; n < p.n; n += (int64_t)gridDim.x * NT) {   // ghost rows untouched
    const int64_t idx = (int64_t)j * p.ldn + n;
    float zz[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      zz[g] = p.P[(int64_t)g * total + idx];
      p.z[g][idx] = zz[g];
      bmax = fmaxf(bmax, fabsf(zz[g]));
    }
    const float i = l_sig(zz[0]), f = l_sig(zz[1]), gg = tanhf(zz[2]), o = l_sig(zz[3]);
    const float c = f * p.c_prev[idx] + i * gg;
    const float h = o * tanhf(c);
    p.gate[0][idx] = i; p.gate[1][idx] = f; p.gate[2][idx] = gg; p.gate[3][idx] = o;
    p.gate[4][idx] = c; p.gate[5][idx] = h;
    l_store_h_side(p, idx, h);
    bmax = fmaxf(bmax, fabsf(h));
    gmax[0] = fmaxf(gmax[0], fabsf(i)); gmax[1] = fmaxf(gmax[1], fabsf(f));       // lambda_p = 0 at the start
    gmax[2] = fmaxf(gmax[2], fabsf(gg)); gmax[3] = fmaxf(gmax[3], fabsf(o));
  }
  track_bound(p.bound_track, bmax);
#pragma unroll
  for (int q = 0; q < 4; ++q) track_max(p.next_max + q, gmax[q]);
}

// ---------------------------------------------------------------------------------------------------- packing
// rows [0,4H): V_g = z_g + lambda_s,g / rho_s ; rows [4H,5H): h_{t-1}; each with its tf32 low part.
struct LPack {
  int64_t n, ldn;
  int32_t H, tc, t0;      // timesteps t0 .. t0+tc-1 (0-based) = slots t0+1 ..
  const float* z[4];      // full tensors [T+1][H][ldn]
  const float* lam_s[4];
  const float* h;
  float rho_s;
  float* r;               // [5H][tc][ldn]
  float* r_lo;
  __half* r16_hi;         // fp16-pair variant (tensor-core path): rows scaled by 2^cap(*r_bound)
  __half* r16_lo;
  const unsigned* r_bound;
};
// exponent c with m 2^c < 2^13 (the same rule as cap_exp() of gate_gemm_tc.cu: producer and consumer must agree)
__device__ __forceinline__ int l_cap_exp(unsigned max_bits) {
  const float m = __uint_as_float(max_bits);
  if (!(m > 0.f) || !isfinite(m)) return 0;
  int e;
  frexpf(m, &e);
  return 13 - e;
}
// grid.x walks the samples two at a time (float2), grid.y the (timestep, unit) rows: no per-element division
__global__ void __launch_bounds__(NT) l_pack_kernel(const LPack p) {
  const int64_t per_t = (int64_t)p.H * p.ldn;
  const float r_scale = p.r16_hi ? ldexpf(1.0f, l_cap_exp(*p.r_bound)) : 1.0f;
  const float irho_s = 1.0f / p.rho_s;
  const int rows = p.H * p.tc;
  for (int row = blockIdx.y; row < rows; row += gridDim.y) {
    const int tl = row / p.H, j = row - tl * p.H;
    for (int64_t n = ((int64_t)blockIdx.x * NT + threadIdx.x) * 2; n < p.ldn; n += (int64_t)gridDim.x * NT * 2) {
      const int64_t rem = (int64_t)j * p.ldn + n;
      const int64_t so = (int64_t)(p.t0 + tl + 1) * per_t + rem;         // slot of timestep t
      const int64_t sp = (int64_t)(p.t0 + tl) * per_t + rem;             // slot of h_{t-1}
      // all loads first: the buffers are not declared restrict, so a store in between would serialise the round trips
      float2 v[5];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float2 z = *reinterpret_cast<const float2*>(p.z[g] + so);
        const float2 l = *reinterpret_cast<const float2*>(p.lam_s[g] + so);
        v[g].x = z.x + div_rn(l.x, p.rho_s, irho_s);
        v[g].y = z.y + div_rn(l.y, p.rho_s, irho_s);
      }
      v[4] = *reinterpret_cast<const float2*>(p.h + sp);
      const bool ok0 = n < p.n, ok1 = n + 1 < p.n;
#pragma unroll
      for (int g = 0; g < 5; ++g) {
        if (!ok0) v[g].x = 0.f;
        if (!ok1) v[g].y = 0.f;
        const int64_t ro = (((int64_t)g * p.H + j) * p.tc + tl) * p.ldn + n;
        if (p.r16_hi) {
          const float c0 = fminf(fmaxf(v[g].x * r_scale, -65504.0f), 65504.0f);
          const float c1 = fminf(fmaxf(v[g].y * r_scale, -65504.0f), 65504.0f);
          const __half h0 = __float2half_rn(c0), h1 = __float2half_rn(c1);
          *reinterpret_cast<__half2*>(p.r16_hi + ro) = __halves2half2(h0, h1);
          *reinterpret_cast<__half2*>(p.r16_lo + ro) =
              __halves2half2(__float2half_rn(c0 - __half2float(h0)), __float2half_rn(c1 - __half2float(h1)));
        } else {
          *reinterpret_cast<float2*>(p.r + ro) = v[g];
          *reinterpret_cast<float2*>(p.r_lo + ro) = make_float2(tf32_lo(v[g].x), tf32_lo(v[g].y));
        }
      }
    }
  }
}

// p_t[j] += sum_n h_T[j][n] (a[n] + lambda11[n]/rho11)     (eq1_W, admm_lstm.py:76-81, the a + lambda/rho part)
__global__ void __launch_bounds__(NT) l_pt_kernel(const float* hT, const float* a, const float* lam11, float rho11,
                                                  int64_t n, int64_t ldn, double* p_t) {
  __shared__ double red[NT / 32];
  const int j = blockIdx.x;
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += NT) acc += (double)hT[(int64_t)j * ldn + i] * (double)(a[i] + lam11[i] / rho11);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < NT / 32; ++w) s += red[w];
    atomicAdd(p_t + j, s);
  }
}

// ---------------------------------------------------------------------------------------------------- sweep
// update_z / update_zg (admm_lstm.py:166-188)
// Every division by a kernel-uniform constant (rho_s, rho_p, rho9, rho10 and the denominators built from them) is
// div_rn(a, b, 1 / b) (admm_math.cuh): the IEEE quotient in three instructions instead of ten -- these kernels execute
// ~600 instructions per element, a third of them such divisions.
struct LRcp {
  float rs, rp, r9, r10;          // rho_s, rho_p, rho9, rho10
  float irs, irp, ir9, ir10;      // their correctly rounded reciprocals
  __device__ __forceinline__ explicit LRcp(const admm_l_hyper& hp)
      : rs(hp.rho_s), rp(hp.rho_p), r9(hp.rho9), r10(hp.rho10), irs(1.0f / hp.rho_s), irp(1.0f / hp.rho_p), ir9(1.0f / hp.rho9),
        ir10(1.0f / hp.rho10) {}
};
struct LZDen {                     // 2 rho_s + rho_p appro of update_z / update_zg and its reciprocal
  float den, iden;
  __device__ __forceinline__ LZDen(const LRcp& k, float appro) : den(2.0f * k.rs + k.rp * appro), iden(1.0f / (2.0f * k.rs + k.rp * appro)) {}
};
__device__ __forceinline__ float l_update_z(float z, float out, float P, float lam1, float lam2, const LRcp& k, float appro,
                                            const LZDen& zd, bool is_g) {
  const float rs = k.rs, rp = k.rp;
  const float act = is_g ? tanhf(z) : l_sig(z);
  const float der = is_g ? 1.0f - act * act : act * (1.0f - act);
  const float form1 = P - div_rn(lam1, rs, k.irs);
  const float form2 = rp * (act - out + div_rn(lam2, rp, k.irp)) * der;
  const float form3 = rs * form1 + 0.5f * rp * appro * z - form2;
  return div_rn(2.0f * form3, zd.den, zd.iden);
}

// z_f,f, z_i,i, z_o,o, z_g,g (main.py:150-165) + the reductions update_c needs
__global__ void __launch_bounds__(NT, 4) l_gates_kernel(const LSlot p, float* red_max, double* red_sum) {
  __shared__ float red[NT / 32];
  __shared__ double redd[NT / 32];
  const int64_t total = (int64_t)p.H * p.ldn;
  const LRcp k(p.hp);
  const float rp = k.rp, r9 = k.r9, r10 = k.r10;
  const float ap_i = appro_sig(red_max[0]), ap_f = appro_sig(red_max[1]), ap_g = appro_tanh(red_max[2]),
              ap_o = appro_sig(red_max[3]);
  const LZDen zd_i(k, ap_i), zd_f(k, ap_f), zd_g(k, ap_g), zd_o(k, ap_o);
  float mx = 0.f, so2 = 0.f;
  for (int j = blockIdx.y; j < p.H; j += gridDim.y)
  for (int64_t n = (int64_t)blockIdx.x * NT + threadIdx.x; n < p.n; n += (int64_t)gridDim.x * NT) {   // ghost rows untouched
    const int64_t idx = (int64_t)j * p.ldn + n;
    const float Pi = p.P[idx], Pf = p.P[total + idx], Pg = p.P[2 * total + idx], Po = p.P[3 * total + idx];
    float i = p.gate[0][idx], f = p.gate[1][idx], g = p.gate[2][idx], o = p.gate[3][idx];
    const float ct = p.gate[4][idx], h = p.gate[5][idx], c_ = p.c_prev[idx];
    const float l9 = p.lam9[idx], l10 = p.lam10[idx];
    const float lpi = p.lam_p[0][idx], lpf = p.lam_p[1][idx], lpg = p.lam_p[2][idx], lpo = p.lam_p[3][idx];
    // f
    const float l9r = div_rn(l9, r9, k.ir9), l10r = div_rn(l10, r10, k.ir10);
    const float zf = l_update_z(p.z[1][idx], f, Pf, p.lam_s[1][idx], lpf, k, ap_f, zd_f, false);
    f = (rp * (l_sig(zf) + div_rn(lpf, rp, k.irp)) + r9 * c_ * (ct - g * i + l9r)) / (rp + r9 * c_ * c_);   // :191-196
    // i
    const float zi = l_update_z(p.z[0][idx], i, Pi, p.lam_s[0][idx], lpi, k, ap_i, zd_i, false);
    i = (rp * (l_sig(zi) + div_rn(lpi, rp, k.irp)) + r9 * g * (ct - c_ * f + l9r)) / (rp + r9 * g * g);     // :199-204
    // o
    const float zo = l_update_z(p.z[3][idx], o, Po, p.lam_s[3][idx], lpo, k, ap_o, zd_o, false);
    const float tc = tanhf(ct);
    o = (rp * (l_sig(zo) + div_rn(lpo, rp, k.irp)) + r10 * tc * (h - l10r)) / (rp + r10 * tc * tc);         // :207-212
    // g
    const float zg = l_update_z(p.z[2][idx], g, Pg, p.lam_s[2][idx], lpg, k, ap_g, zd_g, true);
    g = (rp * (tanhf(zg) + div_rn(lpg, rp, k.irp)) + r9 * i * (ct - c_ * f + l9r)) / (rp + r9 * i * i);     // :215-220
    p.z[0][idx] = zi; p.z[1][idx] = zf; p.z[2][idx] = zg; p.z[3][idx] = zo;
    p.gate[0][idx] = i; p.gate[1][idx] = f; p.gate[2][idx] = g; p.gate[3][idx] = o;
    mx = fmaxf(mx, fabsf((h - l10r) * 1.0f / o));                                                       // :225
    so2 = fmaf(o, o, so2);                                                                              // :230
  }
  const float bm = block_max(mx, red);
  double s = warp_sum((double)so2);
  if ((threadIdx.x & 31) == 0) redd[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    atomic_max_nonneg(red_max + 4, bm);
    double t = 0.0;
    for (int w = 0; w < NT / 32; ++w) t += redd[w];
    atomicAdd(red_sum, t);
  }
}

// The ten dual updates of one element (admm_lstm.py:274-311, main.py:170-188).  Loads, arithmetic and stores are kept
// in three groups: the buffers are not restrict-qualified, so a store between two loads would serialise the memory
// round trips (measured: 0.77 ms instead of ~0.3 ms per timestep at N = 8192, H = 1024).
struct LDualIn {
  float i, f, g, o, l9, l10, z[4], lp[4], ls[4], P[4];
};
__device__ __forceinline__ LDualIn l_duals_load(const LSlot& p, int64_t idx, int64_t total) {
  LDualIn d;
  d.i = p.gate[0][idx]; d.f = p.gate[1][idx]; d.g = p.gate[2][idx]; d.o = p.gate[3][idx];
  d.l9 = p.lam9[idx]; d.l10 = p.lam10[idx];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    d.z[q] = p.z[q][idx]; d.lp[q] = p.lam_p[q][idx]; d.ls[q] = p.lam_s[q][idx]; d.P[q] = p.P[(int64_t)q * total + idx];
  }
  return d;
}
__device__ __forceinline__ void l_duals_store(const LSlot& p, int64_t idx, const LDualIn& d, float c, float h, float c_,
                                              float (&acc)[5], const LRcp& k) {   // acc[0..3]: next iteration's maxima, acc[4]: |V|, |h| bound
  const float rs = k.rs, rp = k.rp, r9 = k.r9, r10 = k.r10;
  const float n10 = d.l10 + r10 * (tanhf(c) * d.o - h);
  const float n9 = d.l9 + r9 * (c - d.g * d.i - c_ * d.f);
  const float gv[4] = {d.i, d.f, d.g, d.o};
  float np_[4], ns_[4];
  float b = fabsf(h);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float act = (q == 2) ? tanhf(d.z[q]) : l_sig(d.z[q]);
    np_[q] = d.lp[q] + rp * (act - gv[q]);
    ns_[q] = d.ls[q] + rs * (d.z[q] - d.P[q]);
    b = fmaxf(b, fabsf(d.z[q] + div_rn(ns_[q], rs, k.irs)));          // |V| of the next packing pass (l_pack_kernel)
    acc[q] = fmaxf(acc[q], fabsf(gv[q] - div_rn(np_[q], rp, k.irp)));  // admm_lstm.py:168,179 of the next iteration
  }
  p.lam10[idx] = n10;
  p.lam9[idx] = n9;
#pragma unroll
  for (int q = 0; q < 4; ++q) { p.lam_p[q][idx] = np_[q]; p.lam_s[q][idx] = ns_[q]; }
  acc[4] = fmaxf(acc[4], b);
}

// update_c (:223-241), update_h for s < T (:249-250), then the duals.  LAST: c only.
template <bool LAST>
__global__ void __launch_bounds__(NT) l_cell_kernel(const LSlot p, const float* red_max, const double* red_sum) {
  const int64_t total = (int64_t)p.H * p.ldn;
  const LRcp k(p.hp);
  const float r9 = k.r9, r10 = k.r10;
  const float appro_h = appro_tanh(red_max[4]);
  const float qua_o = (float)red_sum[0];
  const float form4 = r9 + 0.5f * r10 * qua_o * appro_h;         // uniform: the same for every element of the timestep
  const float iform4 = 1.0f / form4;
  float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int j = blockIdx.y; j < p.H; j += gridDim.y)
  for (int64_t n = (int64_t)blockIdx.x * NT + threadIdx.x; n < p.n; n += (int64_t)gridDim.x * NT) {   // ghost rows untouched
    const int64_t idx = (int64_t)j * p.ldn + n;
    const LDualIn d = l_duals_load(p, idx, total);
    const float i = d.i, f = d.f, g = d.g, o = d.o, l9 = d.l9, l10 = d.l10;
    const float ct = p.gate[4][idx], h = p.gate[5][idx], c_ = p.c_prev[idx];
    const float l10r = div_rn(l10, r10, k.ir10);
    const float form1 = r9 * (g * i + c_ * f - div_rn(l9, r9, k.ir9));
    const float tc = tanhf(ct);
    const float form2 = r10 * (tc * o - h + l10r) * (1.0f - tc * tc) * o;
    const float form3 = 0.5f * r10 * qua_o * ct * appro_h;
    const float c = div_rn(form1 - form2 + form3, form4, iform4);
    if (LAST) {
      p.gate[4][idx] = c;
    } else {
      const float hn = div_rn(r10 * (tanhf(c) * o + l10r), r10, k.ir10);
      p.gate[4][idx] = c;
      p.gate[5][idx] = hn;
      l_store_h_side(p, idx, hn);
      l_duals_store(p, idx, d, c, hn, c_, acc, k);
    }
  }
  if (!LAST) {
    track_bound(p.bound_track, acc[4]);
#pragma unroll
    for (int q = 0; q < 4; ++q) track_max(p.next_max + q, acc[q]);
  }
}

// ---------------------------------------------------------------------------------------------------- t = T-1
// tmp[n] = Form10 = -a + h Wy - lambda11/rho11     (update_h, :252)
__global__ void __launch_bounds__(NT) l_form10_kernel(const float* h, const float* wy, const float* a, const float* lam11,
                                                      float rho11, int64_t n, int64_t ldn, int H, float* tmp) {
  const int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x;
  if (i >= n) return;
  float d = 0.f;
  for (int j = 0; j < H; ++j) d = fmaf(h[(int64_t)j * ldn + i], wy[j], d);
  tmp[i] = -a[i] + d - lam11[i] / rho11;
}
// h_T = (Form1 - rho11 Form11 + theta h)/(rho10 + theta)   (:258)
__global__ void __launch_bounds__(NT) l_last_h_kernel(const LSlot p, const float* wy, const float* tmp, const float* theta_h) {
  const float th = theta_h[0], r10 = p.hp.rho10, r11 = p.hp.rho11;
  for (int j = blockIdx.y; j < p.H; j += gridDim.y)
  for (int64_t n = (int64_t)blockIdx.x * NT + threadIdx.x; n < p.n; n += (int64_t)gridDim.x * NT) {   // ghost rows untouched
    const int64_t idx = (int64_t)j * p.ldn + n;
    const float form1 = r10 * (tanhf(p.gate[4][idx]) * p.gate[3][idx] + p.lam10[idx] / r10);
    const float form11 = tmp[n] * wy[j];
    const float hn = (form1 - r11 * form11 + th * p.gate[5][idx]) / (r10 + th);
    p.gate[5][idx] = hn;
    l_store_h_side(p, idx, hn);
  }
}
// a (:262-266) and lambda11 (:269-272)
__global__ void __launch_bounds__(NT) l_last_a_kernel(const float* h, const float* wy, const float* y, float* a, float* lam11,
                                                      float rho11, float n_norm, float a_den, int64_t n, int64_t ldn, int H) {
  const int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x;
  if (i >= n) return;
  float d = 0.f;
  for (int j = 0; j < H; ++j) d = fmaf(h[(int64_t)j * ldn + i], wy[j], d);
  const float l = lam11[i];
  const float an = (2.0f * y[i] / n_norm + rho11 * d - l) / a_den;
  a[i] = an;
  lam11[i] = l + rho11 * (an - d);
}
__global__ void __launch_bounds__(NT) l_duals_kernel(const LSlot p) {
  const int64_t total = (int64_t)p.H * p.ldn;
  const LRcp k(p.hp);
  float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int j = blockIdx.y; j < p.H; j += gridDim.y)
  for (int64_t n = (int64_t)blockIdx.x * NT + threadIdx.x; n < p.n; n += (int64_t)gridDim.x * NT) {   // ghost rows untouched
    const int64_t idx = (int64_t)j * p.ldn + n;
    const LDualIn d = l_duals_load(p, idx, total);
    l_duals_store(p, idx, d, p.gate[4][idx], p.gate[5][idx], p.c_prev[idx], acc, k);
  }
  track_bound(p.bound_track, acc[4]);
#pragma unroll
  for (int q = 0; q < 4; ++q) track_max(p.next_max + q, acc[q]);
}

// ---------------------------------------------------------------------------------------------------- host helpers
// 2-D grid of ~148*8 CTAs: x over the samples (two per thread), y over rows
dim3 ew_grid2(int64_t ldn, int rows) {
  const int64_t gx = (ldn / 2 + NT - 1) / NT;
  int64_t gy = (148 * 8 + gx - 1) / gx;
  if (gy > rows) gy = rows;
  if (gy < 1) gy = 1;
  return dim3((unsigned)gx, (unsigned)gy);
}

// The gate pass (33 one-float streams per element, bound by memory latency) runs ONE resident wave instead: SMs x the CTAs
// that fit on an SM (64 registers under __launch_bounds__(NT, 4): 4, not the 8 the fixed grids assume -- 1216 CTAs on 444
// slots were 2.74 waves of long-running CTAs), x over sample blocks (the kernel strides over them), y over rows: 52.5 ->
// 45.4 ms per iteration at the configs[4] shape (same-box A/B).  The cell and packing passes measured 3-4 % SLOWER with
// one resident wave of 3 CTAs per SM and keep the fixed grid.
dim3 ew_grid_resident(const void* kernel, int64_t n_blocks, int rows) {
  static const void* known[8];
  static int occ_of[8];
  static int n_known = 0;
  int occ = 0;
  for (int i = 0; i < n_known; ++i)
    if (known[i] == kernel) occ = occ_of[i];
  if (occ == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, NT, 0) != cudaSuccess || occ < 1) occ = 2;
    if (n_known < 8) { known[n_known] = kernel; occ_of[n_known++] = occ; }
  }
  const int64_t slots = 148LL * occ;
  int64_t gx = n_blocks < slots ? n_blocks : slots;
  if (gx < 1) gx = 1;
  int64_t gy = slots / gx;
  if (gy > rows) gy = rows;
  if (gy < 1) gy = 1;
  return dim3((unsigned)gx, (unsigned)gy);
}

// 2-D grid of ~148*8 CTAs for the light kernels: x over the samples (one per thread), y over rows
dim3 ew_grid1(int64_t ldn, int rows) {
  const int64_t gx = (ldn + NT - 1) / NT;
  int64_t gy = (148 * 8 + gx - 1) / gx;
  if (gy > rows) gy = rows;
  if (gy < 1) gy = 1;
  return dim3((unsigned)gx, (unsigned)gy);
}

int validate_l(const admm_l_problem* lp, const char* who) {
  if (!lp) { set_error("%s: null problem", who); return ADMM_EINVAL; }
  int rc = validate(&lp->base, who);
  if (rc) return rc;
  ADMM_REQUIRE(lp->base.O == 1, "%s: ADMM-LSTM-L has a single output (W_y is [H,1], main.py:83)", who);
  for (int g = 0; g < 4; ++g)
    ADMM_REQUIRE(lp->z[g] && lp->lam_s[g] && lp->lam_p[g], "%s: null z / lambda buffers", who);
  ADMM_REQUIRE(lp->lam9 && lp->lam10 && lp->base.a && lp->base.dual_y, "%s: null lambda9 / lambda10 / a / lambda11", who);
  return ADMM_OK;
}

LSlot make_slot(const admm_l_problem* lp, int s, const float* P, float* next_max = nullptr) {
  LSlot k;
  const admm_problem& b = lp->base;
  const int64_t slab = (int64_t)b.H * b.ldn;
  k.n = b.n; k.ldn = b.ldn; k.H = b.H;
  for (int q = 0; q < 6; ++q) k.gate[q] = b.gate[q] + (int64_t)s * slab;
  k.c_prev = b.gate[4] + (int64_t)(s - 1) * slab;
  for (int q = 0; q < 4; ++q) {
    k.z[q] = lp->z[q] + (int64_t)s * slab;
    k.lam_s[q] = lp->lam_s[q] + (int64_t)s * slab;
    k.lam_p[q] = lp->lam_p[q] + (int64_t)s * slab;
  }
  k.lam9 = lp->lam9 + (int64_t)s * slab;
  k.lam10 = lp->lam10 + (int64_t)s * slab;
  k.P = P;
  k.next_max = next_max;
  k.h16_hi = k.h16_lo = nullptr; k.bound_track = nullptr; k.h_ovf = nullptr;
  if (b.tc_ws && tc_eligible(&b)) {
    k.bound_track = tc_r_bound(&b);
    k.h_ovf = tc_h_overflow(&b);
    tc_h16(&b, &k.h16_hi, &k.h16_lo);
    k.h16_hi += (int64_t)s * slab;
    k.h16_lo += (int64_t)s * slab;
  }
  k.hp = lp->hp;
  return k;
}

int reset_bound(const admm_l_problem* lp, cudaStream_t st) {
  const admm_problem& b = lp->base;
  if (!(b.tc_ws && tc_eligible(&b))) return ADMM_OK;
  if (cudaMemsetAsync(tc_r_bound(&b), 0, sizeof(unsigned), st) != cudaSuccess) return check_launch("bound memset");
  return ADMM_OK;
}

int gemm_P(const admm_l_problem* lp, int s, float* scratch, cudaStream_t st) {
  GateGemmArgs a = base_args(&lp->base, s);
  a.scratch = scratch; a.tc = 1;
  return run_gate_gemm(GG_RAWZ, &lp->base, a, 1, st);
}

}  // namespace
}  // namespace admm

using namespace admm;

extern "C" {

int admm_l_sizeof_problem(void) { return (int)sizeof(admm_l_problem); }

int admm_l_forward_t(const admm_l_problem* lp, int s, float* scratch, float* next_max, void* stream) {
  int rc = validate_l(lp, "admm_l_forward_t");
  if (rc) return rc;
  ADMM_REQUIRE(s >= 1 && s <= lp->base.T && scratch && next_max, "admm_l_forward_t: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (s == 1 && (rc = reset_bound(lp, st))) return rc;
  if ((rc = gemm_P(lp, s, scratch, st))) return rc;
  const LSlot k = make_slot(lp, s, scratch, next_max);
  KernelScope ks_("l_forward_kernel", st);
  l_forward_kernel<<<ew_grid1(k.ldn, k.H), NT, 0, st>>>(k);
  count_launch();
  return check_launch("l_forward");
}

int admm_l_output(const admm_l_problem* lp, void* stream) {
  int rc = validate_l(lp, "admm_l_output");
  if (rc) return rc;
  const admm_problem& b = lp->base;
  return launch_output(b.gate[5] + (int64_t)b.T * b.H * b.ldn, b.wy, b.a, b.ldn, b.H, 1, (cudaStream_t)stream);
}

int admm_l_sums(const admm_l_problem* lp, int t0, int tc, float* scratch, double* acc_x, double* acc_h, void* stream) {
  int rc = validate_l(lp, "admm_l_sums");
  if (rc) return rc;
  const admm_problem& b = lp->base;
  ADMM_REQUIRE(t0 >= 0 && tc >= 1 && t0 + tc <= b.T, "admm_l_sums: bad timestep range %d+%d", t0, tc);
  ADMM_REQUIRE(scratch && acc_x && acc_h, "admm_l_sums: null buffers");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t half = 5LL * b.H * tc * b.ldn;
  LPack k;
  k.n = b.n; k.ldn = b.ldn; k.H = b.H; k.tc = tc; k.t0 = t0;
  for (int g = 0; g < 4; ++g) { k.z[g] = lp->z[g]; k.lam_s[g] = lp->lam_s[g]; }
  k.h = b.gate[5]; k.rho_s = lp->hp.rho_s; k.r = scratch; k.r_lo = scratch + half;
  const bool use_tc = b.tc_ws && tc_eligible(&b);
  // fp16 pairs when both reduction GEMMs can take the tensor-core path (row count a multiple of the 128-row tile, D >= 8)
  const bool f16 = use_tc && (5 * b.H) % 128 == 0 && b.D >= 8;
  k.r16_hi = k.r16_lo = nullptr; k.r_bound = nullptr;
  if (f16) {
    k.r16_hi = reinterpret_cast<__half*>(scratch);
    k.r16_lo = reinterpret_cast<__half*>(scratch + half);
    k.r_bound = tc_r_bound(&b);
  }
  {
    KernelScope ks_("l_pack_kernel", st);
    l_pack_kernel<<<ew_grid2(b.ldn, b.H * tc), NT, 0, st>>>(k);
  }
  count_launch();
  if ((rc = check_launch("l_pack"))) return rc;
  AtrArgs r;
  memset(&r, 0, sizeof(r));
  r.ldn = b.ldn; r.H = b.H; r.tc = tc; r.rows = 5 * b.H; r.rpg = b.H;
  r.scratch = scratch; r.scratch_lo = nullptr;            // without fp16 pairs: the fp32 CUDA-core reduction
  if (f16) { r.r16_hi = k.r16_hi; r.r16_lo = k.r16_lo; r.r_bound = k.r_bound; }
  // src = x: P_x[g] and S_xh
  r.K = b.D; r.a_src = b.x + (int64_t)t0 * b.D * b.ldn; r.a_tstride = (int64_t)b.D * b.ldn; r.g_acc = acc_x;
  rc = f16 ? atr_tc(&b, r, st) : atr_simt(r, st);
  if (rc) return rc;
  // src = h_{t-1}: P_h[g] and S_hh
  r.K = b.H; r.a_src = b.gate[5] + (int64_t)t0 * b.H * b.ldn; r.a_tstride = (int64_t)b.H * b.ldn; r.g_acc = acc_h;
  return f16 ? atr_tc(&b, r, st) : atr_simt(r, st);
}

int admm_l_gram_xx(const admm_l_problem* lp, double* sxx, void* stream) {
  int rc = validate_l(lp, "admm_l_gram_xx");
  if (rc) return rc;
  const admm_problem& b = lp->base;
  ADMM_REQUIRE(sxx, "admm_l_gram_xx: null buffer");
  // x is [T][D][ldn]: as R it has D rows per timestep, i.e. T launches with tc = 1 (done once per run)
  for (int t = 0; t < b.T; ++t) {
    AtrArgs r;
    memset(&r, 0, sizeof(r));
    r.ldn = b.ldn; r.H = b.H; r.tc = 1; r.rows = b.D; r.rpg = b.D;
    r.K = b.D; r.a_src = b.x + (int64_t)t * b.D * b.ldn; r.a_tstride = (int64_t)b.D * b.ldn;
    r.scratch = r.a_src; r.scratch_lo = nullptr; r.g_acc = sxx;
    if ((rc = atr_simt(r, (cudaStream_t)stream))) return rc;
  }
  return ADMM_OK;
}

int admm_l_sums_last(const admm_l_problem* lp, double* s_tt, double* p_t, void* stream) {
  int rc = validate_l(lp, "admm_l_sums_last");
  if (rc) return rc;
  const admm_problem& b = lp->base;
  ADMM_REQUIRE(s_tt && p_t, "admm_l_sums_last: null buffers");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t slab = (int64_t)b.H * b.ldn;
  AtrArgs r;
  memset(&r, 0, sizeof(r));
  r.ldn = b.ldn; r.H = b.H; r.tc = 1; r.rows = b.H; r.rpg = b.H;
  r.K = b.H; r.a_src = b.gate[5] + (int64_t)b.T * slab; r.a_tstride = slab;
  r.scratch = r.a_src; r.scratch_lo = nullptr;
  r.g_acc = s_tt;
  rc = atr_simt(r, st);                 // one [H,H] Gram of a single timestep: CUDA-core reduction
  if (rc) return rc;
  KernelScope ks_("l_pt_kernel", st);
  l_pt_kernel<<<b.H, NT, 0, st>>>(b.gate[5] + (int64_t)b.T * slab, b.a, b.dual_y, lp->hp.rho11, b.n, b.ldn, p_t);
  count_launch();
  return check_launch("l_pt");
}

int admm_l_sweep_gates(const admm_l_problem* lp, int s, float* scratch, float* red_max, double* red_sum, void* stream) {
  int rc = validate_l(lp, "admm_l_sweep_gates");
  if (rc) return rc;
  ADMM_REQUIRE(s >= 1 && s <= lp->base.T && scratch && red_max && red_sum, "admm_l_sweep_gates: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (s == 1 && (rc = reset_bound(lp, st))) return rc;       // the sweep re-measures the bound of the next packing pass
  if ((rc = gemm_P(lp, s, scratch, st))) return rc;
  const LSlot k = make_slot(lp, s, scratch);
  KernelScope ks_("l_gates_kernel", st);
  l_gates_kernel<<<ew_grid_resident((const void*)l_gates_kernel, (k.ldn + NT - 1) / NT, k.H), NT, 0, st>>>(k, red_max, red_sum);
  count_launch();
  return check_launch("l_gates");
}

int admm_l_sweep_cell(const admm_l_problem* lp, int s, const float* scratch, const float* red_max, const double* red_sum,
                      float* next_max, void* stream) {
  int rc = validate_l(lp, "admm_l_sweep_cell");
  if (rc) return rc;
  ADMM_REQUIRE(s >= 1 && s <= lp->base.T && scratch && red_max && red_sum && next_max, "admm_l_sweep_cell: bad arguments");
  const LSlot k = make_slot(lp, s, scratch, next_max);
  const dim3 grid = ew_grid1(k.ldn, k.H);
  KernelScope ks_("l_cell_kernel", (cudaStream_t)stream);
  if (s == lp->base.T) l_cell_kernel<true><<<grid, NT, 0, (cudaStream_t)stream>>>(k, red_max, red_sum);
  else l_cell_kernel<false><<<grid, NT, 0, (cudaStream_t)stream>>>(k, red_max, red_sum);
  count_launch();
  return check_launch("l_cell");
}

int admm_l_last(const admm_l_problem* lp, const float* theta_h, float* tmp, const float* scratch, float* next_max,
                void* stream) {
  int rc = validate_l(lp, "admm_l_last");
  if (rc) return rc;
  ADMM_REQUIRE(theta_h && tmp && scratch && next_max, "admm_l_last: null buffers");
  const admm_problem& b = lp->base;
  cudaStream_t st = (cudaStream_t)stream;
  const LSlot k = make_slot(lp, b.T, scratch, next_max);
  const unsigned nb = (unsigned)((b.n + NT - 1) / NT);
  const dim3 grid = ew_grid1(k.ldn, k.H);
  KernelScope ks_("l_last kernels (form10, last_h, last_a, duals)", st);
  l_form10_kernel<<<nb, NT, 0, st>>>(k.gate[5], b.wy, b.a, b.dual_y, lp->hp.rho11, b.n, b.ldn, b.H, tmp);
  l_last_h_kernel<<<grid, NT, 0, st>>>(k, b.wy, tmp, theta_h);
  const float a_den = (float)(2.0 / (double)lp->hp.n_norm + (double)lp->hp.rho11);
  l_last_a_kernel<<<nb, NT, 0, st>>>(k.gate[5], b.wy, b.y, b.a, b.dual_y, lp->hp.rho11, lp->hp.n_norm, a_den, b.n, b.ldn, b.H);
  l_duals_kernel<<<grid, NT, 0, st>>>(k);
  count_launch(4);
  return check_launch("l_last");
}

}  // extern "C"
