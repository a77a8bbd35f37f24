// gate_gemm.h -- argument block shared by the CUDA-core and tensor-core gate GEMM kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "admm_math.cuh"

namespace admm {

enum GateGemmMode { GG_FORWARD = 0, GG_SWEEP = 1, GG_GRAD = 2, GG_PROBE = 3, GG_RAWZ = 7 /* debug: Z -> scratch */ };

// All pointers are pre-offset to the first timestep of the launch (grid.z index tl = 0);
// x_tstride / s_tstride are the element strides from one timestep slab to the next.
struct GateGemmArgs {
  int64_t n, ldn;
  int32_t D, H;
  const float* x;        // x_t slab                  [D][ldn]
  const float* h_prev;   // h_{t-1} slab              [H][ldn]
  const float* c_prev;   // c_{t-1} slab              [H][ldn]   (FORWARD, SWEEP)
  int64_t x_tstride, s_tstride;
  const float* wx;       // [4][D][H]
  const float* wh;       // [4][H][H]
  float* gate[6];        // i f g o c h slabs at t    [H][ldn]
  float* dual[5];        // lambda_i f g o c at t     [H][ldn]
  const float* dual_h;   // lambda_h (t = T only)     [H][ldn]
  Rho rho;
  int32_t last;          // SWEEP: t == T
  double* metrics;       // SWEEP: [>=3] or nullptr
  float* scratch;        // GRAD: R^T [4][H][tc][ldn]
  int32_t tc;
  double* fw_acc;        // GRAD: [4]
  int32_t src;           // PROBE: ADMM_SRC_*
  const float* grad;     // PROBE: G [4][K][H]
  int32_t k0, ncand;     // PROBE: theta_k = 2^(k0+k), k < ncand
  const int32_t* done;   // PROBE: [4]
  double* fk_acc;        // PROBE: [4][NC] with NC = 8 if ncand <= 8 else ADMM_MAX_CAND
  void* tc_ws;           // tensor-core workspace or nullptr
  float* dbg;            // debug dump buffer (RAWZ) or nullptr
  float* h_lo;           // slab t of the h - tf32(h) side buffer (tensor-core path) or nullptr
};

int gate_gemm_simt(int mode, const GateGemmArgs& a, int tc, cudaStream_t st);

// G_acc[4][K][H] (fp64) += A_src^T R  with R^T in `scratch` ([4H][tc][ldn]); a_src pre-offset to the
// slab of the first timestep, a_tstride = K*ldn.
struct AtrArgs {
  int64_t ldn;
  int32_t K, H, tc;
  const float* a_src;
  int64_t a_tstride;
  const float* scratch;
  double* g_acc;
};
int atr_simt(const AtrArgs& a, cudaStream_t st);

}  // namespace admm
