// gate_gemm_simt.cu -- fp32 CUDA-core version of the "gate GEMM + fused epilogue" family.
//
//   Z[n, (g,j)] = sum_k x_t[k,n] * W_g[k,j] + sum_k h_{t-1}[k,n] * U_g[k,j]        g in {i,f,g,o}
//
// followed, in registers, by one of four epilogues:
//   FORWARD : blocks/lstm.py:77-85      (state initialisation / prediction)
//   SWEEP   : admm.py:345-351 + 504-510 (primal i,f,g,o,c,h then dual ascent at timestep t)
//   GRAD    : admm.py:302-312           (R = (act z - lambda/rho - gate) act'(z) -> scratch, f(w) partial sums)
//   PROBE   : admm.py:316-325           (Z0 = A w + B w' and Q = A_src G -> scratch; probe_eval.cu then sums
//                                        f(w + G/theta_k) for a vector of theta_k from them in one pass)
//
// This is the general path (any D, H) and the numerical reference for the tcgen05 path in
// gate_gemm_tc.cu.  Data layout is feature-major ([.][H][ldn], sample index fastest), so every
// global access below is a coalesced float4 over samples.
#include "common.cuh"
#include "gate_gemm.h"

namespace admm {

namespace {

constexpr int BK = 16;     // k-slab per pipeline stage
constexpr int BJ = 32;     // hidden units per CTA (x 4 gates = 128 accumulator columns)
constexpr int NTHREADS = 256;

template <int MODE, int TM>
struct Smem {
  static constexpr int BM = 16 * TM;
  float a[2][BK][BM];
  float w[2][BK][4][BJ];
  float g[(MODE == GG_PROBE) ? 2 : 1][(MODE == GG_PROBE) ? BK : 1][4][(MODE == GG_PROBE) ? BJ : 1];
  float red[4 * (NTHREADS / 32)];
};

template <int MODE, int TM>
__global__ void __launch_bounds__(NTHREADS, 1)
gate_gemm_simt_kernel(const GateGemmArgs p) {
  using S = Smem<MODE, TM>;
  constexpr int BM = S::BM;
  constexpr int NV = TM / 4;            // float4 groups of samples per thread
  __shared__ __align__(16) S sm;

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t n0 = (int64_t)blockIdx.x * BM;
  const int j0 = blockIdx.y * BJ;
  const int tl = blockIdx.z;
  const int D = p.D, H = p.H;
  const int64_t ldn = p.ldn;
  const int ktot = D + H;
  const int nkt = (ktot + BK - 1) / BK;

  if (MODE == GG_PROBE) {
    const int all_done = p.done[0] && p.done[1] && p.done[2] && p.done[3];
    if (all_done) return;
  }

  const float* xs = p.x + (int64_t)tl * p.x_tstride;        // x_t slab      [D][ldn]
  const float* hs = p.h_prev + (int64_t)tl * p.s_tstride;   // h_{t-1} slab  [H][ldn]

  float acc[4][2][TM];
  float accq[(MODE == GG_PROBE) ? 4 : 1][2][(MODE == GG_PROBE) ? TM : 1];
#pragma unroll
  for (int g = 0; g < 4; ++g)
#pragma unroll
    for (int jj = 0; jj < 2; ++jj)
#pragma unroll
      for (int m = 0; m < TM; ++m) {
        acc[g][jj][m] = 0.f;
        if (MODE == GG_PROBE) accq[g][jj][m] = 0.f;
      }

  constexpr int A_ITERS = (BK * BM / 4) / NTHREADS;   // float4 copies of A per thread per stage
  constexpr int W_ITERS = (BK * 4 * BJ) / NTHREADS;   // scalar loads of W per thread per stage
  static_assert(A_ITERS >= 1, "tile too small");
  float wreg[W_ITERS];
  float greg[(MODE == GG_PROBE) ? W_ITERS : 1];

  const int src_lo = (MODE == GG_PROBE) ? (p.src == ADMM_SRC_X ? 0 : D) : 0;
  const int src_hi = (MODE == GG_PROBE) ? (p.src == ADMM_SRC_X ? D : ktot) : 0;

  auto issue_a = [&](int kt, int buf) {
#pragma unroll
    for (int r = 0; r < A_ITERS; ++r) {
      const int idx = tid + r * NTHREADS;
      const int kr = idx / (BM / 4), c4 = idx % (BM / 4);
      const int k = kt * BK + kr;
      float* dst = &sm.a[buf][kr][c4 * 4];
      if (k < ktot) {
        const float* row = (k < D) ? xs + (int64_t)k * ldn : hs + (int64_t)(k - D) * ldn;
        cp_async16(dst, row + n0 + c4 * 4);
      } else {
        *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    cp_async_commit();
  };
  auto load_w = [&](int kt) {
#pragma unroll
    for (int r = 0; r < W_ITERS; ++r) {
      const int idx = tid + r * NTHREADS;
      const int kr = idx / (4 * BJ), rem = idx % (4 * BJ);
      const int g = rem / BJ, jj = rem % BJ;
      const int k = kt * BK + kr, j = j0 + jj;
      float v = 0.f, q = 0.f;
      if (k < ktot && j < H) {
        v = (k < D) ? p.wx[((int64_t)g * D + k) * H + j] : p.wh[((int64_t)g * H + (k - D)) * H + j];
        if (MODE == GG_PROBE && k >= src_lo && k < src_hi)
          q = p.grad[((int64_t)g * (src_hi - src_lo) + (k - src_lo)) * H + j];
      }
      wreg[r] = v;
      if (MODE == GG_PROBE) greg[r] = q;
    }
  };
  auto store_w = [&](int buf) {
#pragma unroll
    for (int r = 0; r < W_ITERS; ++r) {
      const int idx = tid + r * NTHREADS;
      const int kr = idx / (4 * BJ), rem = idx % (4 * BJ);
      sm.w[buf][kr][rem / BJ][rem % BJ] = wreg[r];
      if (MODE == GG_PROBE) sm.g[buf][kr][rem / BJ][rem % BJ] = greg[r];
    }
  };

  issue_a(0, 0);
  load_w(0);
  store_w(0);
  cp_async_wait<0>();
  __syncthreads();

  for (int kt = 0; kt < nkt; ++kt) {
    const int buf = kt & 1;
    const bool more = kt + 1 < nkt;
    if (more) {
      issue_a(kt + 1, buf ^ 1);
      load_w(kt + 1);
    }
    const bool q_tile = (MODE == GG_PROBE) && (kt * BK < src_hi) && (kt * BK + BK > src_lo);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[TM];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const float4 t4 = *reinterpret_cast<const float4*>(&sm.a[buf][kk][v * 64 + tx * 4]);
        av[v * 4 + 0] = t4.x; av[v * 4 + 1] = t4.y; av[v * 4 + 2] = t4.z; av[v * 4 + 3] = t4.w;
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float2 w2 = *reinterpret_cast<const float2*>(&sm.w[buf][kk][g][ty * 2]);
#pragma unroll
        for (int m = 0; m < TM; ++m) {
          acc[g][0][m] = fmaf(av[m], w2.x, acc[g][0][m]);
          acc[g][1][m] = fmaf(av[m], w2.y, acc[g][1][m]);
        }
      }
      if (MODE == GG_PROBE) {
        if (q_tile) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float2 g2 = *reinterpret_cast<const float2*>(&sm.g[buf][kk][g][ty * 2]);
#pragma unroll
            for (int m = 0; m < TM; ++m) {
              accq[g][0][m] = fmaf(av[m], g2.x, accq[g][0][m]);
              accq[g][1][m] = fmaf(av[m], g2.y, accq[g][1][m]);
            }
          }
        }
      }
    }
    if (more) store_w(buf ^ 1);
    cp_async_wait<0>();
    __syncthreads();
  }

  // ------------------------------------------------------------------ epilogues
  const Rho rho = p.rho;
  const int64_t soff = (int64_t)tl * p.s_tstride;     // slab offset of timestep t inside state tensors
  float msum[4] = {0.f, 0.f, 0.f, 0.f};

#pragma unroll
  for (int jj = 0; jj < 2; ++jj) {
    const int j = j0 + ty * 2 + jj;
    if (j >= H) continue;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int64_t n = n0 + v * 64 + tx * 4;
      const int64_t off = soff + (int64_t)j * ldn + n;
      const bool ok[4] = {n + 0 < p.n, n + 1 < p.n, n + 2 < p.n, n + 3 < p.n};

      if (MODE == GG_FORWARD) {
        const float4 cp4 = ld_stream(p.c_prev + off);
        const float cpv[4] = {cp4.x, cp4.y, cp4.z, cp4.w};
        float o_[6][4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const ForwardResult r = forward_point(acc[0][jj][v * 4 + e], acc[1][jj][v * 4 + e],
                                                acc[2][jj][v * 4 + e], acc[3][jj][v * 4 + e], cpv[e]);
          o_[0][e] = r.i; o_[1][e] = r.f; o_[2][e] = r.g; o_[3][e] = r.o; o_[4][e] = r.c; o_[5][e] = r.h;
        }
#pragma unroll
        for (int q = 0; q < 6; ++q)
          if (p.gate[q]) st_stream(p.gate[q] + off, make_float4(o_[q][0], o_[q][1], o_[q][2], o_[q][3]));
      }

      if (MODE == GG_RAWZ) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float* dst = p.scratch + (((int64_t)g * H + j) * p.tc + tl) * ldn + n;
          st_stream(dst, make_float4(acc[g][jj][v * 4 + 0], acc[g][jj][v * 4 + 1], acc[g][jj][v * 4 + 2], acc[g][jj][v * 4 + 3]));
        }
      }

      if (MODE == GG_SWEEP) {
        float in_[13][4];
        {
          const float* srcs[12] = {p.gate[0] + off, p.gate[1] + off, p.gate[2] + off, p.gate[3] + off,
                                   p.gate[4] + off, p.gate[5] + off, p.c_prev + off,  p.dual[0] + off,
                                   p.dual[1] + off, p.dual[2] + off, p.dual[3] + off, p.dual[4] + off};
#pragma unroll
          for (int q = 0; q < 12; ++q) {
            const float4 t4 = ld_stream(srcs[q]);
            in_[q][0] = t4.x; in_[q][1] = t4.y; in_[q][2] = t4.z; in_[q][3] = t4.w;
          }
          float4 lh4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.last) lh4 = ld_stream(p.dual_h + (int64_t)j * ldn + n);
          in_[12][0] = lh4.x; in_[12][1] = lh4.y; in_[12][2] = lh4.z; in_[12][3] = lh4.w;
        }
        float out_[11][4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          SweepPoint s;
          s.zi = acc[0][jj][v * 4 + e]; s.zf = acc[1][jj][v * 4 + e];
          s.zg = acc[2][jj][v * 4 + e]; s.zo = acc[3][jj][v * 4 + e];
          s.i = in_[0][e]; s.f = in_[1][e]; s.g = in_[2][e]; s.o = in_[3][e]; s.c = in_[4][e]; s.h = in_[5][e];
          s.c_prev = in_[6][e];
          s.li = in_[7][e]; s.lf = in_[8][e]; s.lg = in_[9][e]; s.lo = in_[10][e]; s.lc = in_[11][e];
          s.lh = in_[12][e];
          const SweepResult r = sweep_point(s, rho, p.last != 0);
          out_[0][e] = r.i; out_[1][e] = r.f; out_[2][e] = r.g; out_[3][e] = r.o; out_[4][e] = r.c;
          out_[5][e] = r.h;
          out_[6][e] = r.li; out_[7][e] = r.lf; out_[8][e] = r.lg; out_[9][e] = r.lo; out_[10][e] = r.lc;
          if (ok[e]) { msum[0] += r.prim_sq; msum[1] += r.dual_sq; msum[2] += r.penalty; }
        }
#pragma unroll
        for (int q = 0; q < 5; ++q)
          st_stream(p.gate[q] + off, make_float4(out_[q][0], out_[q][1], out_[q][2], out_[q][3]));
        if (!p.last) {
          st_stream(p.gate[5] + off, make_float4(out_[5][0], out_[5][1], out_[5][2], out_[5][3]));
        }
#pragma unroll
        for (int q = 0; q < 5; ++q)
          st_stream(p.dual[q] + off, make_float4(out_[6 + q][0], out_[6 + q][1], out_[6 + q][2], out_[6 + q][3]));
      }

      if (MODE == GG_GRAD) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4 lam4 = ld_stream(p.dual[g] + off);
          const float4 gv4 = ld_stream(p.gate[g] + off);
          const float lam[4] = {lam4.x, lam4.y, lam4.z, lam4.w};
          const float gv[4] = {gv4.x, gv4.y, gv4.z, gv4.w};
          const float rg = (g == 0) ? rho.i : (g == 1) ? rho.f : (g == 2) ? rho.g : rho.o;
          float r_[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float u;
            const float rr = grad_point(acc[g][jj][v * 4 + e], lam[e], gv[e], rg, g == 2, &u);
            r_[e] = ok[e] ? rr : 0.f;
            if (ok[e]) msum[g] += u * u;
          }
          float* dst = p.scratch + (((int64_t)g * H + j) * p.tc + tl) * ldn + n;
          st_stream(dst, make_float4(r_[0], r_[1], r_[2], r_[3]));
        }
      }

      if (MODE == GG_PROBE) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int64_t so = (((int64_t)g * H + j) * p.tc + tl) * ldn + n;
          st_stream(p.scratch + so, make_float4(acc[g][jj][v * 4 + 0], acc[g][jj][v * 4 + 1], acc[g][jj][v * 4 + 2], acc[g][jj][v * 4 + 3]));
          st_stream(p.scratch_q + so, make_float4(accq[g][jj][v * 4 + 0], accq[g][jj][v * 4 + 1], accq[g][jj][v * 4 + 2], accq[g][jj][v * 4 + 3]));
        }
      }
    }
  }

  if (MODE == GG_SWEEP) {
    if (p.metrics) {
      float m3[3] = {msum[0], msum[1], msum[2]};
      block_accumulate<3>(m3, sm.red, p.metrics);
    }
  }
  if (MODE == GG_GRAD) {
    float m4[4] = {msum[0], msum[1], msum[2], msum[3]};
    block_accumulate<4>(m4, sm.red, p.fw_acc);
  }
}

template <int MODE, int TM>
int launch(const GateGemmArgs& a, int tc, cudaStream_t st) {
  constexpr int BM = 16 * TM;
  // all ldn rows, ghosts included: the scratch rows the reduction GEMM reads must all be written (ghost R = 0)
  dim3 grid((unsigned)(a.ldn / BM), (unsigned)((a.H + BJ - 1) / BJ), (unsigned)tc);
  KernelScope ks_("gate_gemm_simt_kernel", st);
  gate_gemm_simt_kernel<MODE, TM><<<grid, NTHREADS, 0, st>>>(a);
  count_launch();
  return check_launch("gate_gemm_simt");
}

}  // namespace

int gate_gemm_simt(int mode, const GateGemmArgs& a, int tc, cudaStream_t st) {
  // Small shards get the 64-sample tile so that more CTAs are in flight.
  const bool small = a.n * ((a.H + BJ - 1) / BJ) * tc < (int64_t)128 * 148 * 2;
  switch (mode) {
    case GG_FORWARD: return small ? launch<GG_FORWARD, 4>(a, tc, st) : launch<GG_FORWARD, 8>(a, tc, st);
    case GG_SWEEP:   return small ? launch<GG_SWEEP, 4>(a, tc, st) : launch<GG_SWEEP, 8>(a, tc, st);
    case GG_GRAD:    return small ? launch<GG_GRAD, 4>(a, tc, st) : launch<GG_GRAD, 8>(a, tc, st);
    case GG_RAWZ:    return launch<GG_RAWZ, 4>(a, tc, st);
    case GG_PROBE:   return launch<GG_PROBE, 4>(a, tc, st);
  }
  set_error("gate_gemm_simt: bad mode %d", mode);
  return ADMM_EINVAL;
}

}  // namespace admm
