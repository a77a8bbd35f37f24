// small_kernels.cu -- the parts of the step that are not gate GEMMs:
//   * Wy update                         (admm.py:246-280 / admm.no_dual_y.py:226-249)
//   * h_T backtracking, a, lambda_h     (admm.py:439-502, 532-546 / admm.no_dual_y.py:414-456)
//   * theta selection + prox for W, U   (admm.py:327-343)
// The N-sized ones are thread-per-sample (feature-major rows make the j loops coalesced); the
// decisions the reference takes in Python `while` loops are replayed on the device from reduced
// fp64 sums, so a step needs no host synchronisation.
#include "common.cuh"
#include "small_kernels.h"

namespace admm {
namespace {

constexpr int NT = 128;

// ------------------------------------------------------------------------------------ Wy
// g_acc[j][o] += sum_n h_T[j][n] * (h_T Wy - a - shift)[n][o]
__global__ void __launch_bounds__(NT) wy_grad_kernel(const admm_problem p, double* g_acc) {
  __shared__ float r_s[ADMM_MAX_O][NT];
  const int64_t n = (int64_t)blockIdx.x * NT + threadIdx.x;
  const int H = p.H, O = p.O;
  const int64_t ldn = p.ldn;
  const float* hT = p.gate[5] + (int64_t)p.T * H * ldn;
  const bool ok = n < p.n;
  float r[ADMM_MAX_O];
#pragma unroll
  for (int o = 0; o < ADMM_MAX_O; ++o) r[o] = 0.f;
  if (ok) {
    for (int j = 0; j < H; ++j) {
      const float hv = hT[(int64_t)j * ldn + n];
#pragma unroll
      for (int o = 0; o < ADMM_MAX_O; ++o)
        if (o < O) r[o] = fmaf(hv, p.wy[j * O + o], r[o]);
    }
#pragma unroll
    for (int o = 0; o < ADMM_MAX_O; ++o)
      if (o < O) {
        r[o] -= p.a[(int64_t)o * ldn + n];
        if (p.with_dual_y) r[o] -= p.dual_y[(int64_t)o * ldn + n] / p.hp.rho[6];
      }
  }
#pragma unroll
  for (int o = 0; o < ADMM_MAX_O; ++o)
    if (o < O) r_s[o][threadIdx.x] = ok ? r[o] : 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t nb = (int64_t)blockIdx.x * NT;
  for (int j = warp; j < H; j += NT / 32) {
    float hv[NT / 32];
#pragma unroll
    for (int q = 0; q < NT / 32; ++q) hv[q] = hT[(int64_t)j * ldn + nb + q * 32 + lane];   // ghosts masked by r_s = 0
    for (int o = 0; o < O; ++o) {
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < NT / 32; ++q) s = fmaf(hv[q], r_s[o][q * 32 + lane], s);
      s = warp_sum(s);
      if (lane == 0) atomicAdd(g_acc + j * O + o, (double)s);
    }
  }
}

__global__ void wy_apply_kernel(const admm_problem p, const double* g_acc) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.H * p.O) return;
  const float rho_y = p.hp.rho[6];
  const float g = rho_y * (float)g_acc[idx];
  // admm.py:277-280 theta = 1 -> 0.5 ; admm.no_dual_y.py:231,247-249 theta = 0.01 -> 0.005, 2*beta
  const float theta = (p.variant == ADMM_VARIANT_ADMM) ? 0.5f : 0.005f;
  const float den = (p.variant == ADMM_VARIANT_ADMM) ? theta + p.hp.beta_wy : theta + 2.0f * p.hp.beta_wy;
  p.wy[idx] = (theta * p.wy[idx] - g) / den;
}

// ------------------------------------------------------------------------------------ t = T
struct LastCtx {
  const float *h, *o, *c, *lh;
  int64_t ldn;
};

__device__ __forceinline__ LastCtx last_ctx(const admm_problem& p) {
  const int64_t slab = (int64_t)p.T * p.H * p.ldn;
  LastCtx c;
  c.h = p.gate[5] + slab; c.o = p.gate[3] + slab; c.c = p.gate[4] + slab; c.lh = p.dual_h;
  c.ldn = p.ldn;
  return c;
}

// r[o] = (h Wy - a - shift)[o] for this sample
__device__ __forceinline__ void residual_y(const admm_problem& p, const LastCtx& c, int64_t n, float* r) {
  const int O = p.O;
#pragma unroll
  for (int o = 0; o < ADMM_MAX_O; ++o) r[o] = 0.f;
  for (int j = 0; j < p.H; ++j) {
    const float hv = c.h[(int64_t)j * c.ldn + n];
#pragma unroll
    for (int o = 0; o < ADMM_MAX_O; ++o)
      if (o < O) r[o] = fmaf(hv, p.wy[j * O + o], r[o]);
  }
#pragma unroll
  for (int o = 0; o < ADMM_MAX_O; ++o)
    if (o < O) {
      r[o] -= p.a[(int64_t)o * c.ldn + n];
      if (p.with_dual_y) r[o] -= p.dual_y[(int64_t)o * c.ldn + n] / p.hp.rho[6];
    }
}

// grad_j for this sample: admm: ((rho_y r) Wy^T)_j   (admm.py:459-464)  fast: rho_h (r Wy^T)_j  (no_dual_y:426)
__device__ __forceinline__ float grad_h(const admm_problem& p, const float* r, int j) {
  const int O = p.O;
  float s = 0.f;
  if (p.variant == ADMM_VARIANT_ADMM) {
#pragma unroll
    for (int o = 0; o < ADMM_MAX_O; ++o)
      if (o < O) s = fmaf(p.hp.rho[6] * r[o], p.wy[j * O + o], s);
    return s;
  }
#pragma unroll
  for (int o = 0; o < ADMM_MAX_O; ++o)
    if (o < O) s = fmaf(r[o], p.wy[j * O + o], s);
  return p.hp.rho[5] * s;
}

constexpr int NTH = 4;   // theta candidates tested by the loop at admm.py:475-480: .1 .2 .4 .8

__global__ void __launch_bounds__(NT) last_probe_kernel(const admm_problem p, double* sums) {
  __shared__ float red[(1 + 3 * NTH) * (NT / 32)];
  const int64_t n = (int64_t)blockIdx.x * NT + threadIdx.x;
  const LastCtx c = last_ctx(p);
  const int O = p.O;
  const float rho_h = p.hp.rho[5];
  float out[1 + 3 * NTH];
#pragma unroll
  for (int k = 0; k < 1 + 3 * NTH; ++k) out[k] = 0.f;
  if (n < p.n) {
    float r[ADMM_MAX_O];
    residual_y(p, c, n, r);
    float bw[NTH][ADMM_MAX_O];
#pragma unroll
    for (int k = 0; k < NTH; ++k)
#pragma unroll
      for (int o = 0; o < ADMM_MAX_O; ++o) bw[k][o] = 0.f;
    float s2[NTH] = {0.f, 0.f, 0.f, 0.f}, s3[NTH] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < p.H; ++j) {
      const int64_t off = (int64_t)j * c.ldn + n;
      const float hv = c.h[off];
      const float g = grad_h(p, r, j);
      const float base = rho_h * c.o[off] * tanh_f(c.c[off]) - c.lh[off];
      float theta = 0.1f;
#pragma unroll
      for (int k = 0; k < NTH; ++k) {
        const float beta = (p.variant == ADMM_VARIANT_ADMM) ? (theta * hv + base - g) / (theta + rho_h)
                                                            : g / theta;
        const float d = beta - hv;
        s2[k] = fmaf(g, d, s2[k]);
        s3[k] = fmaf(d, d, s3[k]);
#pragma unroll
        for (int o = 0; o < ADMM_MAX_O; ++o)
          if (o < O) bw[k][o] = fmaf(beta, p.wy[j * O + o], bw[k][o]);
        theta *= 2.0f;
      }
    }
#pragma unroll
    for (int o = 0; o < ADMM_MAX_O; ++o)
      if (o < O) {
        out[0] = fmaf(r[o], r[o], out[0]);
        const float ao = p.a[(int64_t)o * c.ldn + n] +
                         (p.with_dual_y ? p.dual_y[(int64_t)o * c.ldn + n] / p.hp.rho[6] : 0.f);
#pragma unroll
        for (int k = 0; k < NTH; ++k) {
          const float e = bw[k][o] - ao;
          out[1 + 3 * k] = fmaf(e, e, out[1 + 3 * k]);
        }
      }
#pragma unroll
    for (int k = 0; k < NTH; ++k) { out[2 + 3 * k] = s2[k]; out[3 + 3 * k] = s3[k]; }
  }
  block_accumulate<1 + 3 * NTH>(out, red, sums);
}

// Replays admm.py:472-482 from the reduced sums.
__global__ void last_select_kernel(const admm_problem p, const double* sums, float* theta_out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float rho_y = p.hp.rho[6];
  const float f_h = 0.5f * rho_y * (float)sums[0];
  float theta = 0.1f;
  for (int k = 0; k < NTH; ++k) {
    const float f_b = 0.5f * rho_y * (float)sums[1 + 3 * k];
    const float est = f_h + (float)sums[2 + 3 * k] + (0.5f * theta) * (float)sums[3 + 3 * k];
    if (!(f_b > est)) break;
    theta *= 2.0f;          // after the 4th doubling theta = 1.6 >= theta_max -> break (admm.py:479)
  }
  theta_out[0] = theta * 0.5f;
}

__global__ void __launch_bounds__(NT) last_apply_kernel(const admm_problem p, const float* theta_p, double* metrics) {
  __shared__ float red[4 * (NT / 32)];
  const int64_t n = (int64_t)blockIdx.x * NT + threadIdx.x;
  const LastCtx c = last_ctx(p);
  const int O = p.O;
  const float rho_h = p.hp.rho[5], rho_y = p.hp.rho[6];
  const float theta = theta_p[0];
  float m[4] = {0.f, 0.f, 0.f, 0.f};   // prim_sq, dual_sq, penalty, loss_sq
  if (n < p.ldn) {
    const bool ok = n < p.n;
    float r[ADMM_MAX_O], hw[ADMM_MAX_O];
    residual_y(p, c, n, r);
#pragma unroll
    for (int o = 0; o < ADMM_MAX_O; ++o) hw[o] = 0.f;
    float* hT = p.gate[5] + (int64_t)p.T * p.H * p.ldn;
    for (int j = 0; j < p.H; ++j) {
      const int64_t off = (int64_t)j * c.ldn + n;
      const float hv = c.h[off], lh = c.lh[off];
      const float g = grad_h(p, r, j);
      const float otc = c.o[off] * tanh_f(c.c[off]);
      const float h_new = (theta * hv + rho_h * otc - lh - g) / (theta + rho_h);      // admm.py:483-487
      const float rh = h_new - otc;
      const float lh_new = lh + rho_h * rh;                                           // admm.py:535-539
      hT[off] = h_new;
      p.dual_h[off] = lh_new;
#pragma unroll
      for (int o = 0; o < ADMM_MAX_O; ++o)
        if (o < O) hw[o] = fmaf(h_new, p.wy[j * O + o], hw[o]);
      if (ok) {
        const float dh = h_new - hv;
        m[0] = fmaf(rh, rh, m[0]);
        m[1] = fmaf(rho_h * rho_h * dh, dh, m[1]);
        m[2] += lh_new * rh + 0.5f * rho_h * rh * rh;
      }
    }
    const float nb = (float)p.n_global;
#pragma unroll
    for (int o = 0; o < ADMM_MAX_O; ++o)
      if (o < O) {
        const int64_t off = (int64_t)o * c.ldn + n;
        const float a_old = p.a[off], yv = p.y[off];
        float num = 2.0f * yv + nb * rho_y * hw[o];                                   // admm.py:496-501
        if (p.with_dual_y) num -= nb * p.dual_y[off];
        const float a_new = num / (2.0f + nb * rho_y);
        p.a[off] = a_new;
        const float ry = a_new - hw[o];
        float pen = 0.5f * rho_y * ry * ry;
        if (p.with_dual_y) {
          const float ly = p.dual_y[off] + rho_y * ry;                                // admm.py:541-546
          p.dual_y[off] = ly;
          pen += ly * ry;
        }
        if (ok) {
          const float da = a_new - a_old, e = a_new - yv;
          m[0] = fmaf(ry, ry, m[0]);
          m[1] = fmaf(rho_y * rho_y * da, da, m[1]);
          m[2] += pen;
          m[3] = fmaf(e, e, m[3]);
        }
      }
  }
  if (metrics) block_accumulate<4>(m, red, metrics);
}

// a = h_T Wy for the forward initialisation (blocks/lstm.py:87)
__global__ void __launch_bounds__(NT) output_kernel(const float* hT, const float* wy, float* a, int64_t ldn,
                                                    int H, int O) {
  const int64_t n = (int64_t)blockIdx.x * NT + threadIdx.x;
  if (n >= ldn) return;
  float acc[ADMM_MAX_O];
#pragma unroll
  for (int o = 0; o < ADMM_MAX_O; ++o) acc[o] = 0.f;
  for (int j = 0; j < H; ++j) {
    const float hv = hT[(int64_t)j * ldn + n];
#pragma unroll
    for (int o = 0; o < ADMM_MAX_O; ++o)
      if (o < O) acc[o] = fmaf(hv, wy[j * O + o], acc[o]);
  }
#pragma unroll
  for (int o = 0; o < ADMM_MAX_O; ++o)
    if (o < O) a[(int64_t)o * ldn + n] = acc[o];
}

// ------------------------------------------------------------------------------------ W / U
__global__ void weight_finish_kernel(const double* g_acc, float* grad, const admm_hyper hp, int64_t per_gate) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 4 * per_gate) return;
  const int g = (int)(idx / per_gate);
  grad[idx] = (float)g_acc[idx] * hp.rho[g];          // admm.py:312
}

// est() ingredients for every theta = 2^k, k < ADMM_EST_CAND, once per weight update (w and G do not change between
// probe passes):  s1_k = <G, beta_k - w>, s2_k = ||beta_k - w||^2 with beta_k = fl(w + G/2^k) rounded to fp32 exactly
// as the reference forms it (admm.py:332,336).  grid = (4 gates, ADMM_EST_CAND/8 candidate groups, element blocks).
__global__ void __launch_bounds__(256) weight_est_kernel(const float* w_all, const float* grad_all, int64_t per_gate,
                                                         double* est_acc) {
  __shared__ double red[16][8];
  const int g = blockIdx.x, kg = blockIdx.y * 8;
  const float* w = w_all + (int64_t)g * per_gate;
  const float* G = grad_all + (int64_t)g * per_gate;
  double s1[8], s2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { s1[k] = 0.0; s2[k] = 0.0; }
  for (int64_t e = (int64_t)blockIdx.z * blockDim.x + threadIdx.x; e < per_gate; e += (int64_t)gridDim.z * blockDim.x) {
    const float wv = w[e], gv = G[e];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float beta = wv + gv * ldexpf(1.0f, -(kg + k));      // G / 2^k is exact
      const float d = beta - wv;
      s1[k] += (double)(gv * d);
      s2[k] += (double)(d * d);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const double a = warp_sum(s1[k]), b = warp_sum(s2[k]);
    if (lane == 0) { red[k][warp] = a; red[8 + k][warp] = b; }
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    double s = 0.0;
    for (int q = 0; q < 8; ++q) s += red[threadIdx.x][q];
    const int k = kg + (threadIdx.x & 7), which = threadIdx.x >> 3;
    atomicAdd(est_acc + ((int64_t)g * ADMM_EST_CAND + k) * 2 + which, s);
  }
}

// Replays admm.py:331-338 per gate from the reduced sums: theta_k = 2^k, est_k = f(w) + s1_k + T*0.5*theta_k*s2_k, exit at
// the first k with !(f(beta_k) > est_k).  With plan.proof the candidates below the window are represented by
// lower bounds of f(beta_k): "bound > est_k" proves the loop continues; an unproven k leaves the gate undecided.
__global__ void weight_select_kernel(const double* est_acc, const double* fk_acc, const float* qmax, admm_hyper hp, int T,
                                     admm_probe_plan plan, int final_pass, int32_t* done, float* theta_out) {
  const int g = threadIdx.x;
  if (g >= 4 || done[g]) return;
  const float rho = hp.rho[g];
  const double* fk = fk_acc + g * ADMM_FK_SLOTS;
  const double* es = est_acc + (int64_t)g * ADMM_EST_CAND * 2;
  const float f_w = 0.5f * rho * (float)fk[plan.moments ? ADMM_FK_MOMENTS : ADMM_MAX_CAND];
  auto est = [&](int k) {
    const float theta = ldexpf(1.0f, k);
    return f_w + (float)es[2 * k] + ((float)T * 0.5f * theta) * (float)es[2 * k + 1];
  };
  const int k0 = plan.k0[g];
  if (plan.moments) {
    // f(w + G/2^k) = F0 + B1 x + ... + B6 x^6, x = 2^-k, valid where max|Q| x <= 2^-4 (admm_probe_plan).  kv = first
    // exponent at which it is; every k < kv must be covered by a lower bound that proves the loop continues past it.
    int kv = k0;
    const float valid = (plan.order == 4) ? 0.015625f : 0.0625f;       // |Q| 2^-k <= 2^-6 (order 4) / 2^-4 (order 6)
    while (kv < ADMM_EST_CAND && !(ldexpf(qmax[g], -kv) <= valid)) ++kv;
    const int kp = plan.proof ? k0 + plan.ncand : 0;
    if (kv > kp && kv > 0) {
      done[4 + g] = 3; done[8 + g] = kv;             // expansion not valid where the bounds end: exact passes follow
      return;
    }
    for (int k = 0; k < kv; ++k) {
      const float lower = 0.5f * rho * (float)fk[ADMM_MAX_CAND + 1 + k];
      if (!(lower > est(k))) {
        done[4 + g] = 1; done[8 + g] = k;
        return;
      }
    }
    const double* m = fk + ADMM_FK_MOMENTS;
    for (int k = kv; k < ADMM_EST_CAND; ++k) {
      const double x = ldexp(1.0, -k);
      const double f_k = m[0] + x * (m[1] + x * (m[2] + x * (m[3] + x * (m[4] + x * (m[5] + x * m[6])))));
      const float f_b = 0.5f * rho * (float)f_k;
      if (!(f_b > est(k))) {
        theta_out[g] = ldexpf(1.0f, k - 1);          // theta /= 2 (admm.py:338)
        done[g] = 1;
        return;
      }
    }
    theta_out[g] = ldexpf(1.0f, ADMM_EST_CAND - 1);   // iteration cap (the reference has none)
    done[g] = 1;
    return;
  }
  if (plan.proof) {
    for (int k = 0; k < k0; ++k) {
      const float lower = 0.5f * rho * (float)fk[ADMM_MAX_CAND + 1 + k];
      if (!(lower > est(k))) {                       // not provable from the subsample: a full pass will decide
        done[4 + g] = 1; done[8 + g] = k;
        return;
      }
    }
  }
  int found = -1;
  for (int c = 0; c < plan.ncand; ++c) {
    const int k = k0 + c;
    if (k >= ADMM_EST_CAND) break;
    const float f_b = 0.5f * rho * (float)fk[c];
    if (!(f_b > est(k))) { found = k; break; }
  }
  if (found >= 0) {
    theta_out[g] = ldexpf(1.0f, found - 1);            // theta /= 2 (admm.py:338)
    done[g] = 1;
  } else if (!final_pass) {
    done[4 + g] = 2; done[8 + g] = k0 + plan.ncand;    // diagnostics: window exhausted
  } else if (final_pass) {
    theta_out[g] = ldexpf(1.0f, k0 + plan.ncand - 1);  // iteration cap (the reference has none)
    done[g] = 1;
  }
}

__global__ void weight_apply_kernel(float* w_all, const float* grad_all, const float* theta, admm_hyper hp,
                                    int src, int T, int64_t per_gate) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 4 * per_gate) return;
  const int g = (int)(idx / per_gate);
  const float rho = hp.rho[g];
  const float beta = (src == ADMM_SRC_X) ? hp.beta_x[g] : hp.beta_h[g];
  const float th = theta[g];
  const float tau_n = 0.5f * rho * (float)T * th;          // admm.py:341, left to right
  const float tau_d = 0.5f * rho * th * (float)T;          // admm.py:342
  w_all[idx] = (tau_n * w_all[idx] - grad_all[idx]) / (beta + tau_d);
}

}  // namespace

// ------------------------------------------------------------------------------------ launchers
int launch_wy_grad(const admm_problem& p, double* g_acc, cudaStream_t st) {
  const unsigned blocks = (unsigned)(p.ldn / NT);
  KernelScope ks_("wy_grad_kernel", st);
  wy_grad_kernel<<<blocks, NT, 0, st>>>(p, g_acc);
  count_launch();
  return check_launch("wy_grad");
}
int launch_wy_apply(const admm_problem& p, const double* g_acc, cudaStream_t st) {
  const int tot = p.H * p.O;
  KernelScope ks_("wy_apply_kernel", st);
  wy_apply_kernel<<<(tot + 255) / 256, 256, 0, st>>>(p, g_acc);
  count_launch();
  return check_launch("wy_apply");
}
int launch_last_probe(const admm_problem& p, double* sums, cudaStream_t st) {
  KernelScope ks_("last_probe_kernel", st);
  last_probe_kernel<<<(unsigned)(p.ldn / NT), NT, 0, st>>>(p, sums);
  count_launch();
  return check_launch("last_probe");
}
int launch_last_select(const admm_problem& p, const double* sums, float* theta, cudaStream_t st) {
  KernelScope ks_("last_select_kernel", st);
  last_select_kernel<<<1, 32, 0, st>>>(p, sums, theta);
  count_launch();
  return check_launch("last_select");
}
int launch_last_apply(const admm_problem& p, const float* theta, double* metrics, cudaStream_t st) {
  KernelScope ks_("last_apply_kernel", st);
  last_apply_kernel<<<(unsigned)(p.ldn / NT), NT, 0, st>>>(p, theta, metrics);
  count_launch();
  return check_launch("last_apply");
}
int launch_output(const float* hT, const float* wy, float* a, int64_t ldn, int H, int O, cudaStream_t st) {
  KernelScope ks_("output_kernel", st);
  output_kernel<<<(unsigned)(ldn / NT), NT, 0, st>>>(hT, wy, a, ldn, H, O);
  count_launch();
  return check_launch("output");
}
int launch_weight_finish(const admm_problem& p, int src, const double* g_acc, float* grad, cudaStream_t st) {
  const int64_t per_gate = (int64_t)(src == ADMM_SRC_X ? p.D : p.H) * p.H;
  KernelScope ks_("weight_finish_kernel", st);
  weight_finish_kernel<<<(unsigned)((4 * per_gate + 255) / 256), 256, 0, st>>>(g_acc, grad, p.hp, per_gate);
  count_launch();
  return check_launch("weight_finish");
}
int launch_weight_est(const admm_problem& p, int src, const float* grad, double* est_acc, cudaStream_t st) {
  const int64_t per_gate = (int64_t)(src == ADMM_SRC_X ? p.D : p.H) * p.H;
  const float* w = (src == ADMM_SRC_X) ? p.wx : p.wh;
  if (cudaMemsetAsync(est_acc, 0, sizeof(double) * 4 * ADMM_EST_CAND * 2, st) != cudaSuccess) return check_launch("est memset");
  const unsigned nz = (unsigned)((per_gate + 256 * 64 - 1) / (256 * 64));
  dim3 grid(4, ADMM_EST_CAND / 8, nz < 1 ? 1 : (nz > 32 ? 32 : nz));
  KernelScope ks_("weight_est_kernel", st);
  weight_est_kernel<<<grid, 256, 0, st>>>(w, grad, per_gate, est_acc);
  count_launch();
  return check_launch("weight_est");
}
int launch_weight_select(const admm_problem& p, const double* est_acc, const double* fk_acc, const float* qmax,
                         const admm_probe_plan& plan, int final_pass, int32_t* done, float* theta, cudaStream_t st) {
  KernelScope ks_("weight_select_kernel", st);
  weight_select_kernel<<<1, 32, 0, st>>>(est_acc, fk_acc, qmax, p.hp, p.T, plan, final_pass, done, theta);
  count_launch();
  return check_launch("weight_select");
}
int launch_weight_apply(const admm_problem& p, int src, const float* grad, const float* theta, cudaStream_t st) {
  const int64_t per_gate = (int64_t)(src == ADMM_SRC_X ? p.D : p.H) * p.H;
  float* w = (src == ADMM_SRC_X) ? p.wx : p.wh;
  KernelScope ks_("weight_apply_kernel", st);
  weight_apply_kernel<<<(unsigned)((4 * per_gate + 255) / 256), 256, 0, st>>>(w, grad, theta, p.hp, src, p.T, per_gate);
  count_launch();
  return check_launch("weight_apply");
}

}  // namespace admm
