// admm_math.cuh -- the closed-form per-element updates of the ADMM sweep.
//
// Pure scalar fp32 functions, usable from device code and (for the CPU formula tests in
// tests/test_point_math.py) from a host-only build.  Each cites the reference lines it replaces.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define ADMM_HD __host__ __device__ __forceinline__
#else
#define ADMM_HD static inline
#endif

namespace admm {

struct Rho { float i, f, g, o, c, h, y; };

ADMM_HD float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }     // admm.py:235-237
ADMM_HD float tanh_f(float x) { return tanhf(x); }                         // admm.py:231-233
// admm.py:239-244, expressed through the activation value itself
ADMM_HD float dsigmoid_from(float s) { return s * (1.0f - s); }
ADMM_HD float dtanh_from(float t) { return 1.0f - t * t; }

// admm.py:384-386: -(lam - rho1*act(z) + (rho2*(p2*p3 - var2) - lam2)*p1) / (rho1 + rho2*p1*p1)
ADMM_HD float gate_prox(float lam, float rho1, float act_z, float rho2, float p2p3, float var2,
                        float lam2, float p1) {
  return -(lam - rho1 * act_z + (rho2 * (p2p3 - var2) - lam2) * p1) / (rho1 + rho2 * p1 * p1);
}

struct SweepPoint {
  // in: pre-activations of this (sample, unit) at t, old state at t, c at t-1, duals at t
  float zi, zf, zg, zo;
  float i, f, g, o, c, h, c_prev;
  float li, lf, lg, lo, lc, lh;
};

struct SweepResult {
  float i, f, g, o, c, h;
  float li, lf, lg, lo, lc;
  float prim_sq, dual_sq, penalty;   // metric contributions
};

// One (sample, hidden unit) of admm.py:345-351 followed by admm.py:504-510 at timestep t.
// last == true (t == T): h and lambda_h are NOT touched here (admm_last_* does them); the old h and
// lambda_h still enter the o and c updates exactly as in the reference.
ADMM_HD SweepResult sweep_point(const SweepPoint& s, const Rho& r, bool last) {
  SweepResult out;
  const float ai = sigmoid_f(s.zi), af = sigmoid_f(s.zf), ag = tanh_f(s.zg), ao = sigmoid_f(s.zo);
  // i: p1 = g_t, p2 = f_t, p3 = c_{t-1}   (admm.py:361-364)
  out.i = gate_prox(s.li, r.i, ai, r.c, s.f * s.c_prev, s.c, s.lc, s.g);
  // f: p1 = c_{t-1}, p2 = g_t, p3 = i_t(new)   (admm.py:365-368)
  out.f = gate_prox(s.lf, r.f, af, r.c, s.g * out.i, s.c, s.lc, s.c_prev);
  // g: p1 = i_t(new), p2 = f_t(new), p3 = c_{t-1}   (admm.py:369-372)
  out.g = gate_prox(s.lg, r.g, ag, r.c, out.f * s.c_prev, s.c, s.lc, out.i);
  // o: p1 = tanh(c_t old), p2 = p3 = 0, var2 = h_t old   (admm.py:373-379)
  const float tc_old = tanh_f(s.c);
  out.o = gate_prox(s.lo, r.o, ao, r.h, 0.0f, s.h, s.lh, tc_old);
  // c: admm.py:388-436 with theta = 0.5 (the loop at :430 never iterates)
  const float zed = s.h + s.lh / r.h;
  const float u = tc_old * out.o - zed;
  const float grad = (u * out.o) * (1.0f - tc_old * tc_old);
  const float A = s.lc / r.c - out.f * s.c_prev - out.i * out.g;
  out.c = (0.5f * s.c - grad - r.c * A) / (r.c + 0.5f);
  // h, t < T: admm.py:455-457
  const float tc_new = tanh_f(out.c);
  out.h = last ? s.h : (r.h * out.o * tc_new - s.lh) / r.h;
  // duals: admm.py:512-530 (same z as the primal update: weights and h_{t-1} are unchanged)
  const float ri = out.i - ai, rf = out.f - af, rg = out.g - ag, ro = out.o - ao;
  const float rc = out.c - (out.f * s.c_prev + out.i * out.g);
  out.li = s.li + r.i * ri;
  out.lf = s.lf + r.f * rf;
  out.lg = s.lg + r.g * rg;
  out.lo = s.lo + r.o * ro;
  out.lc = s.lc + r.c * rc;
  out.prim_sq = ri * ri + rf * rf + rg * rg + ro * ro + rc * rc;
  const float di = out.i - s.i, df = out.f - s.f, dg = out.g - s.g, dob = out.o - s.o, dc = out.c - s.c;
  const float dh = out.h - s.h;
  out.dual_sq = r.i * r.i * di * di + r.f * r.f * df * df + r.g * r.g * dg * dg + r.o * r.o * dob * dob +
                r.c * r.c * dc * dc + r.h * r.h * dh * dh;
  out.penalty = out.li * ri + 0.5f * r.i * ri * ri + out.lf * rf + 0.5f * r.f * rf * rf +
                out.lg * rg + 0.5f * r.g * rg * rg + out.lo * ro + 0.5f * r.o * ro * ro +
                out.lc * rc + 0.5f * r.c * rc * rc;
  return out;
}

// blocks/lstm.py:80-85
struct ForwardResult { float i, f, g, o, c, h; };
ADMM_HD ForwardResult forward_point(float zi, float zf, float zg, float zo, float c_prev) {
  ForwardResult out;
  out.i = sigmoid_f(zi);
  out.f = sigmoid_f(zf);
  out.g = tanh_f(zg);
  out.o = sigmoid_f(zo);
  out.c = out.f * c_prev + out.i * out.g;
  out.h = out.o * tanh_f(out.c);
  return out;
}

// admm.py:302-312: residual u = act(z) - lambda/rho - gate; R = u * act'(z).  gate_is_g selects tanh.
ADMM_HD float grad_point(float z, float lam, float gate, float rho, bool gate_is_g, float* u_out) {
  const float a = gate_is_g ? tanh_f(z) : sigmoid_f(z);
  const float d = gate_is_g ? dtanh_from(a) : dsigmoid_from(a);
  const float u = a - lam / rho - gate;
  *u_out = u;
  return u * d;
}

// admm.py:316-325 summand for beta = w + G/theta:  (act(z0 + q/theta) - lambda/rho - gate)^2, with the
// reference's association (act - lambda/rho) - gate: the residual is a cancellation, so the order matters
// for the knife-edge comparisons of the backtracking loop.
ADMM_HD float probe_point(float z0, float q, float inv_theta, float lam_over_rho, float gate, bool gate_is_g) {
  const float z = z0 + q * inv_theta;
  const float a = gate_is_g ? tanh_f(z) : sigmoid_f(z);
  const float u = (a - lam_over_rho) - gate;
  return u * u;
}

}  // namespace admm
