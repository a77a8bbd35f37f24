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

void admm_host_probe_points(const float* z0, const float* q, float inv_theta, const float* lam, const float* gate,
                            float rho, int is_g, float* out, long n) {
  for (long e = 0; e < n; ++e)
    out[e] = admm::probe_point(z0[e], q[e], inv_theta, lam[e] / rho, gate[e], is_g != 0);
}
}
