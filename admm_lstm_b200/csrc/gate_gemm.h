// gate_gemm.h -- argument block shared by the CUDA-core and tensor-core gate GEMM kernels.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "admm_math.cuh"

namespace admm {

enum GateGemmMode { GG_FORWARD = 0, GG_SWEEP = 1, GG_GRAD = 2, GG_PROBE = 3, GG_MOMENTS = 4 /* Q + moment sums, fused */,
                    GG_RAWZ = 7 /* Z -> scratch */ };

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
  float* scratch;        // GRAD: R^T [4][H][tc][ldn] ; PROBE: Z0 in the same layout
  float* scratch_q;      // PROBE: Q = A_src G, same layout ; GRAD (tensor-core path): tf32 low part of R^T
  int32_t tc;
  double* fw_acc;        // GRAD: [4]
  int32_t src;           // PROBE: ADMM_SRC_*
  const float* grad;     // PROBE: G [4][K][H]
  const int32_t* done;   // PROBE: [4], the launch is skipped on the device when all four are set
  void* tc_ws;           // tensor-core workspace or nullptr
  float* zstore;         // [4][H][zT][ldn] pre-activation store (tensor-core path) or nullptr
  int32_t zT, zt0;       // zstore timesteps, first timestep (0-based) of this launch
  int32_t z_accumulate;  // GRAD: z = zstore + acc (acc = x (W_new - W_old)) instead of z = acc
  __half* h16_hi;        // slab t of the fp16 pair of h 2^11 (tensor-core path): the gate GEMM's A operand
  __half* h16_lo;
  unsigned* h_ovf;       // sticky device flag: an |h| >= 2^16 / 2^11 was clamped when its fp16 pair was written (set by gate_gemm_tc)
  const float* acc_scale;  // tensor-core path: 2^-(sa+sb) that turns the accumulator of this launch into z or Q
  // GRAD on the tensor-core path.  x-phase: bound_track receives max (1 + |lambda/rho| + |gate|) >= |R| (bit pattern,
  // atomicMax).  h-phase: R^T is written as an fp16 pair of R 2^sR (sR from *r_bound) for the fp16 A^T R GEMM.
  unsigned* bound_track;
  const unsigned* r_bound;
  __half* r16_hi;
  __half* r16_lo;
  int32_t epi_prefetch;  // tensor-core path, SWEEP: L2 prefetch of the epilogue's next batch of units (set by gate_gemm_tc)
  // SWEEP on the tensor-core path: receives max (1 + |lambda_g/rho_g| + |gate_g|) over the values the sweep WRITES, g = i,f,g,o
  // (bit pattern, atomicMax).  Those are exactly what the next iteration's gradient passes read, so the bound on |R| that
  // scales the fp16 operand of A^T R is known before the x-phase starts and its R^T can be fp16 pairs too.
  unsigned* xbound_track;
  int32_t tma_hint;      // tensor-core path: L2 evict_last hint on the weight-operand TMA loads (set by gate_gemm_tc)
  // Device-side launch predicate (tensor-core path): when set, the kernel returns at once unless *run_if != 0.  The x-phase
  // gradient pass launches BOTH the pass over stored pre-activations (grad_from_z, skip_if) and this GEMM pass (run_if) on
  // the "inputs changed since the z store was written" flag, so re-installing identical inputs costs no host round trip.
  const int32_t* run_if;
  // MOMENTS (tensor-core path with a valid z store): the probe operand Q = A_src G stays in TMEM and the epilogue accumulates
  // the moment sums of admm_probe_plan::moments straight from it (Q never goes to HBM): fk_acc[g][ADMM_FK_MOMENTS + 0..6],
  // qmax[g], and -- on 1/8 of the units, a rigorous lower bound like the subsets of the unfused path -- the exact
  // sums of the candidates k < mom_pc[g] into fk_acc[g][ADMM_MAX_CAND + 1 + k].  mom_k0 = the plan's k0 (normalisation of Q).
  int32_t mom_k0[4];
  int32_t mom_pc[4];     // proof candidates per gate, <= 16
  int32_t mom_order;     // 6 or 4 (admm_probe_plan::order)
  double* fk_acc;
  float* qmax;
};

int gate_gemm_simt(int mode, const GateGemmArgs& a, int tc, cudaStream_t st);

// G_acc[4][K][H] (fp64) += A_src^T R  with R^T in `scratch` ([4H][tc][ldn]); a_src pre-offset to the
// slab of the first timestep, a_tstride = K*ldn.
struct AtrArgs {
  int64_t ldn;
  int32_t K, H, tc;
  const float* a_src;
  int64_t a_tstride;
  const float* scratch;     // R^T
  const float* scratch_lo;  // tf32 low part of R^T (tensor-core path) or nullptr
  double* g_acc;
  // Generalisation used by the Gram / right-hand-side sums of ADMM-LSTM-L: R has `rows` rows (0 = the default 4*H)
  // in groups of `rpg` (0 = H): G_acc[c / rpg][k][c % rpg] += sum A_src[k][n] R[c][n].
  int32_t rows, rpg;
  // fp16-pair variant (tensor-core path): R^T as fp16 pairs of R 2^sR, sR = cap(*r_bound); A_src from the fp16 pairs
  // of x / h kept by the tensor-core workspace.  nullptr -> 3xTF32 on `scratch` / `scratch_lo`.
  const __half* r16_hi;
  const __half* r16_lo;
  const unsigned* r_bound;
};
int atr_simt(const AtrArgs& a, cudaStream_t st);

// fk_acc[4][ADMM_FK_SLOTS] (fp64) += sum over the chunk of (act(Z0 + Q 2^-k) - lambda/rho - gate)^2 for the thetas
// of the plan, and, in slot 32 of each gate, the same sum at Q = 0, i.e. f(w)   (admm.py:316-325).
struct ProbeEvalArgs {
  int64_t n, ldn;
  int32_t H, tc;
  int32_t z_T, z_t0;     // timestep extent / first timestep of z0's layout ([4][H][z_T][ldn]); q is always [4][H][tc][ldn]
  const float* z0;
  const float* q;
  const float* gate[4];  // gate_g slab of the first timestep of the chunk
  const float* dual[4];
  int64_t s_tstride;
  float rho[4];
  int32_t kbase[4];      // per gate: first theta exponent evaluated by this launch
  int32_t nc[4];         // per gate: number of consecutive exponents
  int32_t slot0[4];      // per gate: fk_acc slot of the first one
  int32_t jmod, jrem;    // only unit blocks jb with jb % jmod == jrem are summed (jmod = 1: all units)
  int32_t publish_fw;    // also publish f(w) into slot ADMM_MAX_CAND
  const int32_t* done;
  double* fk_acc;        // [4][ADMM_FK_SLOTS]
};
int probe_eval(const ProbeEvalArgs& a, cudaStream_t st);
// Moment pass of a `moments` plan (admm_probe_plan): fk_acc[g][ADMM_FK_MOMENTS + 0..4] += F0, B1..B4 and
// qmax[g] = max(qmax[g], max |Q|).  Uses z0, q, gate, dual, rho, done, kbase (= the plan's k0: normalisation 2^-k0 of Q)
// and fk_acc of the same argument block.
int probe_moments(const ProbeEvalArgs& a, float* qmax, cudaStream_t st);

// x-phase gradient residual from stored pre-activations (grad_from_z.cu): R^T, its tf32 low part and f(w).
struct GradFromZArgs {
  int64_t n, ldn;
  int32_t H, tc;
  int32_t zT, zt0;        // zstore is [4][H][zT][ldn]; first timestep (0-based) of the chunk
  const float* zstore;
  const float* gate[4];   // gate_g / lambda_g slabs of the first timestep of the chunk
  const float* dual[4];
  int64_t s_tstride;
  float rho[4];
  float* r;               // [4][H][tc][ldn]
  float* r_lo;
  double* fw_acc;         // [4]
  unsigned* bound_track;  // max (1 + |lambda/rho| + |gate|) over the elements read (bit pattern, atomicMax) or nullptr
  // fp16-pair output (when the bound on |R| is already known, GateGemmArgs::xbound_track): R 2^sR as hi / lo halves in the
  // layout of r / r_lo (half the bytes), sR = cap(*r_bound); nullptr -> fp32 R and its tf32 low part
  __half* r16_hi;
  __half* r16_lo;
  const unsigned* r_bound;
  const int32_t* skip_if;   // device flag: return at once when *skip_if != 0 (see GateGemmArgs::run_if)
};
int grad_from_z(const GradFromZArgs& a, cudaStream_t st);

}  // namespace admm
