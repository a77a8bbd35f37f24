// point_math_host.cpp -- host-only wrapper around admm_math.cuh so the CPU test-suite can check
// the per-element closed forms (the same source the CUDA kernels inline) against the oracle.
#include "admm_math.cuh"

extern "C" {

// in: [17][n] rows = zi zf zg zo i f g o c h c_prev li lf lg lo lc lh ; out: [14][n] rows =
// i f g o c h li lf lg lo lc prim_sq dual_sq penalty ; rho: i f g o c h y
void admm_host_sweep_points(const float* in, float* out, long n, const float* rho, int last) {
  admm::Rho r{rho[0], rho[1], rho[2], rho[3], rho[4], rho[5], rho[6]};
  for (long e = 0; e < n; ++e) {
    admm::SweepPoint s;
    s.zi = in[0 * n + e]; s.zf = in[1 * n + e]; s.zg = in[2 * n + e]; s.zo = in[3 * n + e];
    s.i = in[4 * n + e]; s.f = in[5 * n + e]; s.g = in[6 * n + e]; s.o = in[7 * n + e];
    s.c = in[8 * n + e]; s.h = in[9 * n + e]; s.c_prev = in[10 * n + e];
    s.li = in[11 * n + e]; s.lf = in[12 * n + e]; s.lg = in[13 * n + e]; s.lo = in[14 * n + e];
    s.lc = in[15 * n + e]; s.lh = in[16 * n + e];
    const admm::SweepResult q = admm::sweep_point(s, r, last != 0);
    const float v[14] = {q.i, q.f, q.g, q.o, q.c, q.h, q.li, q.lf, q.lg, q.lo, q.lc, q.prim_sq, q.dual_sq, q.penalty};
    for (int k = 0; k < 14; ++k) out[k * n + e] = v[k];
  }
}

// R and u of admm.py:302-312 for one gate
void admm_host_grad_points(const float* z, const float* lam, const float* gate, float rho, int is_g, float* R,
                           float* u, long n) {
  for (long e = 0; e < n; ++e) R[e] = admm::grad_point(z[e], lam[e], gate[e], rho, is_g != 0, &u[e]);
}

// lambda / rho the way the fused moment pass forms it (div_rn from the correctly rounded reciprocal)
void admm_host_div_rn(const float* a, float b, float* out, long n) {
  const float y = 1.0f / b;
  for (long e = 0; e < n; ++e) out[e] = admm::div_rn(a[e], b, y);
}

// sums over the elements of c_k t^k, k = 1..4 (double accumulation of the fp32 terms): [0..3] from moment_terms4 (coefficients,
// then powers), [4..7] from moment_accum4 (the fused form the epilogue runs); s = act(z) is formed here in double precision
void admm_host_moment_sums4(const float* z, const float* c, const float* t, int is_g, double* out, long n) {
  for (int k = 0; k < 8; ++k) out[k] = 0.0;
  for (long e = 0; e < n; ++e) {
    const float s = is_g ? (float)tanh((double)z[e]) : (float)(1.0 / (1.0 + exp(-(double)z[e])));
    const float u = s - c[e];
    float co[4];
    admm::moment_terms4(is_g != 0, s, u, co);
    const float t1 = t[e], t2 = t1 * t1, t3 = t2 * t1;
    out[0] += (double)(co[0] * t1);
    out[1] += (double)(co[1] * t2);
    out[2] += (double)(co[2] * t3);
    out[3] += (double)((co[3] * t2) * t2);
    float a[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    admm::moment_accum4(is_g != 0, s, u, t1, a);
    for (int k = 0; k < 4; ++k) out[4 + k] += (double)a[1 + k];
  }
}

void admm_host_probe_points(const float* z0, const float* q, float inv_theta, const float* lam, const float* gate,
                            float rho, int is_g, float* out, long n) {
  for (long e = 0; e < n; ++e)
    out[e] = admm::probe_point(z0[e], q[e], inv_theta, lam[e] / rho, gate[e], is_g != 0);
}
}
