// admm_math.cuh -- the closed-form per-element updates of the ADMM sweep.
//
// Pure scalar fp32 functions, usable from device code and (for the CPU formula tests in
// tests/test_point_math.py) from a host-only build.  Each cites the reference lines it replaces.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define ADMM_HD __host__ __device__ __forceinline__
#define ADMM_HDM __host__ __device__ __forceinline__
#else
#define ADMM_HD static inline
#define ADMM_HDM inline
#endif

namespace admm {

struct Rho { float i, f, g, o, c, h, y; };

ADMM_HD float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }     // admm.py:235-237
ADMM_HD float tanh_f(float x) { return tanhf(x); }                         // admm.py:231-233

// a / b correctly rounded from y = RN(1 / b) (Markstein: q = RN(a y), r = a - b q exactly by FMA, RN(q + r y) is the
// correctly rounded quotient): the value an IEEE division returns for every normal-range operand, in three instructions
// instead of the ten of the general division (reciprocal refinement, FCHK, slow-path branch).  Used where the reference
// divides by a per-gate constant (lambda / rho, admm.py:318) inside an instruction-bound epilogue; tests/test_cabi_and_host.py
// compares it bit for bit with the host's division.
ADMM_HD float div_rn(float a, float b, float y) {
  const float q = a * y;
  return fmaf(fmaf(-b, q, a), y, q);
}

// Math policies of the per-element closed forms.  Accurate = libm-class functions and IEEE division (the CUDA-core
// path, the host build, everything the knife-edge parity tests look at).  Fast (device only) = two MUFU ops per
// activation (ex2.approx + rcp.approx with one Newton step; degree-5 odd polynomial for |x| < 0.6 in tanh) and
// reciprocal-multiply division, <= 2 ulp: used by the tensor-core epilogues, whose instruction count bounds the kernel.
struct AccurateMath {
  ADMM_HDM static float sigmoid(float x) { return sigmoid_f(x); }
  ADMM_HDM static float tanh(float x) { return tanh_f(x); }
  ADMM_HDM static float div(float a, float b) { return a / b; }
};
#if defined(__CUDACC__)
struct FastMath {
  __device__ __forceinline__ static float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
  }
  // 1/d for normal positive or negative d: rcp.approx + one Newton step (~0.5 ulp)
  __device__ __forceinline__ static float rcp(float d) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    return fmaf(r, fmaf(-d, r, 1.0f), r);
  }
  __device__ __forceinline__ static float sigmoid(float x) {
    return rcp(1.0f + ex2(-1.4426950408889634f * fmaxf(x, -80.0f)));
  }
  __device__ __forceinline__ static float tanh(float x) {
    const float ax = fminf(fabsf(x), 40.0f);
    const float e = ex2(-2.8853900817779268f * ax);
    const float big = (1.0f - e) * rcp(1.0f + e);
    const float t = x * x;
    float p = 0.0024976127315312624f;
    p = fmaf(p, t, -0.008506525307893753f);
    p = fmaf(p, t, 0.02181374467909336f);
    p = fmaf(p, t, -0.05396425351500511f);
    p = fmaf(p, t, 0.133333221077919f);
    p = fmaf(p, t, -0.3333333432674408f);
    const float small = fmaf(x * t, p, x);
    return ax < 0.6f ? small : copysignf(big, x);
  }
  __device__ __forceinline__ static float div(float a, float b) { return a * rcp(b); }
};
#endif
#if defined(__CUDACC__)
// Coefficients of (act(z + d) - c)^2 - (act(z) - c)^2 = c1 d + c2 d^2 + ... + c6 d^6 + O(d^7) with s = act(z),
// u = act(z) - c (the moment pass of the backtracking probes, include/admm_lstm_b200.h admm_probe_plan::moments).
// a_j = act^(j)(z) / j!; for |d| <= 2^-4 the truncated terms are < 2e-10 (tanh) / 1e-12 (sigmoid) -- checked against
// float64 in scripts/ and by test_moment_probe_equals_exact_probe.
__device__ __forceinline__ void moment_terms(bool is_g, float s, float u, float (&c)[6]) {
  float a1, a2, a3, a4, a5, a6;
  if (is_g) {
    // tanh (t = s, D1 = 1 - t^2): D2 = -2 t D1, D3 = -2 D1 (1 - 3 t^2), D4 = 8 t D1 (2 - 3 t^2),
    // D5 = 8 D1 (2 - 15 t^2 + 15 t^4), D6 = -16 t D1 (17 - 60 t^2 + 45 t^4)
    const float t2 = s * s, d1 = 1.0f - t2;
    a1 = d1;
    a2 = -s * d1;
    a3 = (-1.0f / 3.0f) * d1 * fmaf(-3.0f, t2, 1.0f);
    a4 = (1.0f / 3.0f) * s * d1 * fmaf(-3.0f, t2, 2.0f);
    a5 = (1.0f / 15.0f) * d1 * fmaf(fmaf(15.0f, t2, -15.0f), t2, 2.0f);
    a6 = (-1.0f / 45.0f) * s * d1 * fmaf(fmaf(45.0f, t2, -60.0f), t2, 17.0f);
  } else {
    // sigmoid (D1 = s (1 - s), m = 1 - 2 s): D2 = D1 m, D3 = D1 (1 - 6 D1), D4 = D1 m (1 - 12 D1),
    // D5 = D1 (1 - 30 D1 + 120 D1^2), D6 = D1 m (1 - 60 D1 + 360 D1^2)
    const float d1 = s * (1.0f - s), m = fmaf(-2.0f, s, 1.0f), dm = d1 * m;
    a1 = d1;
    a2 = 0.5f * dm;
    a3 = (1.0f / 6.0f) * d1 * fmaf(-6.0f, d1, 1.0f);
    a4 = (1.0f / 24.0f) * dm * fmaf(-12.0f, d1, 1.0f);
    a5 = (1.0f / 120.0f) * d1 * fmaf(fmaf(120.0f, d1, -30.0f), d1, 1.0f);
    a6 = (1.0f / 720.0f) * dm * fmaf(fmaf(360.0f, d1, -60.0f), d1, 1.0f);
  }
  const float u2 = 2.0f * u;
  c[0] = u2 * a1;
  c[1] = fmaf(u2, a2, a1 * a1);
  c[2] = fmaf(u2, a3, 2.0f * a1 * a2);
  c[3] = fmaf(u2, a4, fmaf(2.0f * a1, a3, a2 * a2));
  c[4] = fmaf(u2, a5, 2.0f * fmaf(a1, a4, a2 * a3));
  c[5] = fmaf(u2, a6, fmaf(2.0f, fmaf(a1, a5, a2 * a4), a3 * a3));
}
#endif

// The same up to 4th order (c1..c4): valid for |d| <= 2^-6 at the accuracy moment_terms has for |d| <= 2^-4.
ADMM_HD void moment_terms4(bool is_g, float s, float u, float (&c)[4]) {
  float a1, a2, a3, a4;
  if (is_g) {
    const float t2 = s * s, d1 = 1.0f - t2;
    a1 = d1;
    a2 = -s * d1;
    a3 = (-1.0f / 3.0f) * d1 * fmaf(-3.0f, t2, 1.0f);
    a4 = (1.0f / 3.0f) * s * d1 * fmaf(-3.0f, t2, 2.0f);
  } else {
    const float d1 = s * (1.0f - s), m = fmaf(-2.0f, s, 1.0f), dm = d1 * m;
    a1 = d1;
    a2 = 0.5f * dm;
    a3 = (1.0f / 6.0f) * d1 * fmaf(-6.0f, d1, 1.0f);
    a4 = (1.0f / 24.0f) * dm * fmaf(-12.0f, d1, 1.0f);
  }
  const float u2 = 2.0f * u;
  c[0] = u2 * a1;
  c[1] = fmaf(u2, a2, a1 * a1);
  c[2] = fmaf(u2, a3, 2.0f * a1 * a2);
  c[3] = fmaf(u2, a4, fmaf(2.0f * a1, a3, a2 * a2));
}

// moment_terms4 and its accumulation in one: a[k] += c_k t^k, k = 1..4, with the common factor D1 = act'(z) taken into the
// powers (w_k = D1 t^k) and the brackets of c_k / D1 written out -- 21 (tanh) / 23 (sigmoid) instructions instead of 33:
//   sigmoid (m = 1 - 2 s):  c1 = D1 2u,  c2 = D1 (D1 + u m),  c3 = D1 (D1 m + u (1/3 - 2 D1)),
//                           c4 = D1 (D1 (m^2 / 4 + 1/3 - 2 D1) + u m (1/12 - D1))
//   tanh (e = 1/3 - s^2):   c1 = D1 2u,  c2 = D1 (D1 - 2u s),  c3 = D1 (-2 s D1 - 2u e),
//                           c4 = D1 (D1 (s^2 - 2 e) + 2u s (2/3 - s^2))
// (the same polynomials as moment_terms4, checked term by term in float64 and for bias in fp32 sums).
ADMM_HD void moment_accum4(bool is_g, float s, float u, float t, float* a) {
  const float u2 = u + u;
  if (is_g) {
    const float s2 = s * s, d1 = 1.0f - s2, e = (1.0f / 3.0f) - s2, us = u2 * s;
    const float w1 = d1 * t, w2 = w1 * t, w3 = w2 * t, w4 = w3 * t;
    a[1] = fmaf(w1, u2, a[1]);
    a[2] = fmaf(w2, d1 - us, a[2]);
    a[3] = fmaf(w3, fmaf(-2.0f, s * d1, -(u2 * e)), a[3]);
    a[4] = fmaf(w4, fmaf(d1, fmaf(-2.0f, e, s2), us * ((2.0f / 3.0f) - s2)), a[4]);
  } else {
    const float d1 = s * (1.0f - s), m = fmaf(-2.0f, s, 1.0f), e = fmaf(-2.0f, d1, 1.0f / 3.0f), um = u * m;
    const float w1 = d1 * t, w2 = w1 * t, w3 = w2 * t, w4 = w3 * t;
    a[1] = fmaf(w1, u2, a[1]);
    a[2] = fmaf(w2, d1 + um, a[2]);
    a[3] = fmaf(w3, fmaf(u, e, d1 * m), a[3]);
    a[4] = fmaf(w4, fmaf(d1, fmaf(0.25f, m * m, e), um * ((1.0f / 12.0f) - d1)), a[4]);
  }
}

// admm.py:239-244, expressed through the activation value itself
ADMM_HD float dsigmoid_from(float s) { return s * (1.0f - s); }
ADMM_HD float dtanh_from(float t) { return 1.0f - t * t; }

// admm.py:384-386: -(lam - rho1*act(z) + (rho2*(p2*p3 - var2) - lam2)*p1) / (rho1 + rho2*p1*p1)
template <class M = AccurateMath>
ADMM_HD float gate_prox(float lam, float rho1, float act_z, float rho2, float p2p3, float var2,
                        float lam2, float p1) {
  return M::div(-(lam - rho1 * act_z + (rho2 * (p2p3 - var2) - lam2) * p1), rho1 + rho2 * p1 * p1);
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
template <class M = AccurateMath>
ADMM_HD SweepResult sweep_point(const SweepPoint& s, const Rho& r, bool last) {
  SweepResult out;
  const float ai = M::sigmoid(s.zi), af = M::sigmoid(s.zf), ag = M::tanh(s.zg), ao = M::sigmoid(s.zo);
  // i: p1 = g_t, p2 = f_t, p3 = c_{t-1}   (admm.py:361-364)
  out.i = gate_prox<M>(s.li, r.i, ai, r.c, s.f * s.c_prev, s.c, s.lc, s.g);
  // f: p1 = c_{t-1}, p2 = g_t, p3 = i_t(new)   (admm.py:365-368)
  out.f = gate_prox<M>(s.lf, r.f, af, r.c, s.g * out.i, s.c, s.lc, s.c_prev);
  // g: p1 = i_t(new), p2 = f_t(new), p3 = c_{t-1}   (admm.py:369-372)
  out.g = gate_prox<M>(s.lg, r.g, ag, r.c, out.f * s.c_prev, s.c, s.lc, out.i);
  // o: p1 = tanh(c_t old), p2 = p3 = 0, var2 = h_t old   (admm.py:373-379)
  const float tc_old = M::tanh(s.c);
  out.o = gate_prox<M>(s.lo, r.o, ao, r.h, 0.0f, s.h, s.lh, tc_old);
  // c: admm.py:388-436 with theta = 0.5 (the loop at :430 never iterates)
  const float zed = s.h + M::div(s.lh, r.h);
  const float u = tc_old * out.o - zed;
  const float grad = (u * out.o) * (1.0f - tc_old * tc_old);
  const float A = M::div(s.lc, r.c) - out.f * s.c_prev - out.i * out.g;
  out.c = M::div(0.5f * s.c - grad - r.c * A, r.c + 0.5f);
  // h, t < T: admm.py:455-457
  const float tc_new = M::tanh(out.c);
  out.h = last ? s.h : M::div(r.h * out.o * tc_new - s.lh, r.h);
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
template <class M = AccurateMath>
ADMM_HD ForwardResult forward_point(float zi, float zf, float zg, float zo, float c_prev) {
  ForwardResult out;
  out.i = M::sigmoid(zi);
  out.f = M::sigmoid(zf);
  out.g = M::tanh(zg);
  out.o = M::sigmoid(zo);
  out.c = out.f * c_prev + out.i * out.g;
  out.h = out.o * M::tanh(out.c);
  return out;
}

// admm.py:302-312: residual u = act(z) - lambda/rho - gate; R = u * act'(z).  gate_is_g selects tanh.
template <class M = AccurateMath>
ADMM_HD float grad_point(float z, float lam, float gate, float rho, bool gate_is_g, float* u_out) {
  const float a = gate_is_g ? M::tanh(z) : M::sigmoid(z);
  const float d = gate_is_g ? dtanh_from(a) : dsigmoid_from(a);
  const float u = a - M::div(lam, rho) - gate;
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
