/* admm_lstm_b200.h -- C ABI of the B200-native ADMM-LSTM sweep.
 *
 * Drop-in boundary for the per-iteration ADMM sweep of Frederick2309/ADMM-LSTM
 * (reference: admm.py / admm.no_dual_y.py `ADMMBasedOptimizer.step()`).  The reference has no
 * FFI of its own (it is pure Python/torch); the binding a maintainer adds is the ctypes stub in
 * INTEGRATION.md, and `admm_lstm_b200/optimizer.py` is that stub fleshed out behind the
 * reference's class/method names.  Every entry point replaces one reference function (file:line
 * cited per function).  Plain pointers and sizes only: no torch types cross this line.
 *
 * Conventions
 *  - All pointers are DEVICE pointers unless named host_*.  All arithmetic is IEEE fp32
 *    (the reference's precision); cross-sample sums are accumulated in fp64.
 *  - Every call is asynchronous on `stream` (a cudaStream_t passed as void*).
 *  - Return value: 0 on success, otherwise a negative ADMM_E* code; admm_last_error() gives text.
 *  - Device layout ("feature-major"): a state tensor the reference holds as
 *    [N, T+1, H] lives here as [T+1][H][ldn] (sample index fastest, ldn = N rounded up to 128);
 *    inputs x [N,T,D] as [T][D][ldn]; a, y, lambda_y [N,O] as [O][ldn].  Rows n >= N are ghost
 *    samples: they are updated like real ones but masked out of every sum.
 *  - Weights keep the reference's own layout: x2g [D,H], h2g [H,H] stacked per gate in the order
 *    i,f,g,o -> wx [4][D][H], wh [4][H][H]; Wy = out [H][O].
 */
#ifndef ADMM_LSTM_B200_H
#define ADMM_LSTM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADMM_ABI_VERSION 1

#define ADMM_OK 0
#define ADMM_EINVAL (-1)   /* bad argument / unsupported shape            */
#define ADMM_ECUDA (-2)    /* a CUDA runtime call or kernel launch failed */
#define ADMM_ENODEV (-3)   /* no sm_100 device                            */

#define ADMM_VARIANT_ADMM 0       /* admm.py (shipped with_dual_y = False)   */
#define ADMM_VARIANT_NO_DUAL_Y 1  /* admm.no_dual_y.py, the "Fast" variant   */

#define ADMM_SRC_X 0  /* map_from == 'x' : W  = x2g [D,H] */
#define ADMM_SRC_H 1  /* map_from == 'h' : U  = h2g [H,H] */

#define ADMM_MAX_O 16         /* output_size limit of the t = T kernels */
#define ADMM_MAX_CAND 32      /* theta candidates per probe pass                                   */
#define ADMM_EST_CAND 64      /* est() sums are precomputed for theta = 2^0 .. 2^63                 */
#define ADMM_FK_SLOTS 72      /* [0,32) window candidates, [32] f(w), [33,65) lower-bound sums k<32, [65,72) moments */
#define ADMM_FK_MOMENTS 65    /* F0 = f(w) sum, then B1..B6 of the expansion below                                    */
#define ADMM_N_METRICS 8

/* rho / beta in the reference's own key order (admm.py:131-160, parameters.py). */
typedef struct admm_hyper {
  float rho[7];  /* i f g o c h y */
  float beta_x[4]; /* wi wf wg wo  (x2g) */
  float beta_h[4]; /* vi vf vg vo  (h2g) */
  float beta_wy;
} admm_hyper;

/* Which thetas a probe pass evaluates, per gate (identical on every rank: derived from replicated data).
 * Window: theta = 2^(k0[g]+c), c < ncand, summed over ALL units -> fk_acc[g][c]; f(w) -> fk_acc[g][32].
 * proof != 0: additionally, for k < k0[g], the same sum over a fixed subset of the units (1/8 of them, 5/8 for the
 * three exponents next to the window) -> fk_acc[g][33+k].  A
 * partial sum of squares is a rigorous LOWER bound of f(w + G/2^k); if even the bound exceeds est_k the
 * reference's loop provably continues past k (admm.py:334) without the full evaluation.  If a bound does
 * not prove it, the pass stays undecided for that gate and a full pass from k = 0 follows.
 *
 * moments != 0 (the normal first pass): no candidate is evaluated one by one.  With delta = Q 2^-k the perturbation of the
 * pre-activation, act(z + delta) = act(z) + a1 delta + ... + a6 delta^6 + O(delta^7), so
 *      f(w + G/2^k) - f(w) = sum_j B_j 2^(-jk),   B_j = sum over elements of c_j(z, u) Q^j   (u = act z - lambda/rho - gate,
 *      c_1 = 2 u a1, c_2 = 2 u a2 + a1^2, c_3 = 2 u a3 + 2 a1 a2, ..., c_6 = 2 u a6 + 2 a1 a5 + 2 a2 a4 + a3^2)
 * and ONE pass with ONE activation per element (-> fk_acc[g][65..71] = F0, B1..B6 and qmax[g] = max |Q|) gives f for every
 * k >= k0[g] with max|Q| 2^-k0 <= 2^-4, where the truncated terms are < 2e-10 per element -- below the rounding of the
 * reference's own fp32 evaluation (measured: max |delta| at the exit is 1e-4 .. 1e-7 on the benchmark workloads).
 * Candidates k < k0[g] (perturbation too large for the expansion; usually none) are handled by the lower-bound proofs
 * above; if the expansion is not valid at k0[g] the gate stays undecided and the exact passes follow. */
typedef struct admm_probe_plan {
  int32_t k0[4];
  int32_t ncand;
  int32_t proof;
  int32_t moments;
  /* moments plans: order of the expansion, 6 (0 means 6) or 4.  Order 4 drops the delta^5 and delta^6 terms -- a quarter fewer
   * instructions per element in the kernel that is bound by them -- and is used where max|Q| 2^-k <= 2^-6 (instead of 2^-4)
   * keeps the truncated terms below 1e-10 per element; k0 is then two exponents larger. */
  int32_t order;
} admm_probe_plan;

/* One rank's shard of the problem.  The optimizer owns every buffer (allocated by the host
 * language -- torch here); this struct only borrows them for the duration of a call. */
typedef struct admm_problem {
  int64_t n;        /* samples in this shard                                   */
  int64_t n_global; /* samples over all shards (admm.py:497 `self.batch_size`) */
  int64_t ldn;      /* padded sample stride, multiple of 128, >= n             */
  int32_t T, D, H, O;
  int32_t variant;      /* ADMM_VARIANT_*          */
  int32_t with_dual_y;  /* admm.py:12, default 0   */
  admm_hyper hp;
  const float* x;  /* [T][D][ldn]     */
  const float* y;  /* [O][ldn]        */
  float* gate[6];  /* i f g o c h : each [T+1][H][ldn]             */
  float* dual[5];  /* lambda_i f g o c : each [T+1][H][ldn]        */
  float* dual_h;   /* lambda_h at t = T only : [H][ldn]  (zero for t < T, admm.py:533) */
  float* a;        /* [O][ldn]        */
  float* dual_y;   /* [O][ldn]        */
  float* wx;       /* [4][D][H]       */
  float* wh;       /* [4][H][H]       */
  float* wy;       /* [H][O]          */
  /* Optional operands of the tensor-core path (NULL -> fp32 CUDA-core path), see
   * admm_tc_workspace_bytes(): TF32 hi/lo splits kept by the library.  Must be zero-filled once
   * by the caller before first use. */
  void* tc_ws;
  int64_t tc_ws_bytes;
  /* Optional (tensor-core path only; NULL -> the pre-activations are recomputed in every pass):
   * zstore [4][H][T][ldn] keeps z = x W + h U of all timesteps between the passes of the weight phase, so the
   * phase needs 3 instead of 7 full-size GEMM passes per step; wx_prev [4][D][H] is the copy of
   * x2g taken before its update, from which the h-phase refreshes zstore with x (W_new - W_old) only. */
  float* zstore;
  float* wx_prev;
  /* Set by the caller when zstore holds z = x W + h U of the CURRENT x, h, wx, wh for every timestep -- true after
   * the forward initialisation or a complete sweep t = 1..T (both write zstore from the GEMM they run anyway), false
   * after anything else changed inputs, state or weights.  When set, the x-phase gradient pass of the next step runs
   * no GEMM at all: it is an elementwise pass over zstore. */
  int32_t z_valid;
  int32_t reserved_;
} admm_problem;

/* Slots of the fp64 metric accumulator filled by admm_sweep_t / admm_last_apply. */
enum {
  ADMM_M_PRIMAL_SQ = 0, /* sum ||r||^2 over all constraints                         */
  ADMM_M_DUAL_SQ = 1,   /* sum rho_v^2 ||v_new - v_old||^2 over i,f,g,o,c,h,a       */
  ADMM_M_PENALTY = 2,   /* sum <lambda_new, r> + rho/2 ||r||^2                      */
  ADMM_M_LOSS_SQ = 3    /* ||a - y||^2 (divide by n_global for the loss term)       */
};

const char* admm_last_error(void);
int admm_abi_version(void);
/* sizeof(admm_problem) as compiled, so a foreign-language binding can verify its struct layout. */
int admm_sizeof_problem(void);
/* Number of visible sm_100 devices usable by this library (0 on a CPU-only box). */
int admm_device_ok(void);

/* blocks/lstm.py:65-88 init_gate_variables: writes i,f,g,o,c,h at t (reads h,c at t-1) and, at
 * t == T with a != NULL, a = h_T Wy.  Call for t = 1..T in order. */
int admm_forward_t(const admm_problem* p, int t, void* stream);
/* Prediction only (demo.py:341-342 model(x)): runs all T steps keeping two h/c slabs.
 * work: 4 * H * ldn floats.  out: [O][ldn]. */
int admm_predict(const admm_problem* p, float* work, float* out, void* stream);

/* admm.py:246-280 / admm.no_dual_y.py:226-249  __update_wy.
 * grad: g_acc[H*O] (fp64, zeroed by the caller) += h_T^T (h_T Wy - a [- lambda_y/rho_y]);
 * all-reduce g_acc over shards, then apply: Wy <- (theta Wy - rho_y g)/(theta + c beta_wy). */
int admm_wy_grad(const admm_problem* p, double* g_acc, void* stream);
int admm_wy_apply(const admm_problem* p, const double* g_acc, void* stream);

/* admm.py:282-343 __update_weights, all four gates of one `src` batched (the gates do not
 * couple in the weight phase).
 *  grad   : for timesteps t0 < t <= t0+tc: z = x_t W + h_{t-1} U; R = (act z - lambda/rho - gate) act'(z);
 *           g_acc[4][K][H] (fp64) += A_src^T R;  fw_acc[4] (fp64) += sum (act z - lambda/rho - gate)^2.
 *           scratch: 8*H*tc*ldn floats (R^T and, on the tensor-core path, its tf32 low part).
 *  finish : G = rho_g * (float) g_acc  -> grad_out [4][K][H] fp32           (admm.py:312); also est_acc
 *           [4][ADMM_EST_CAND][2] (fp64) = <G, beta_k - w>, ||beta_k - w||^2 for beta_k = fl(w + G/2^k) (admm.py:327-332)
 *  probe  : fk_acc[4][ADMM_FK_SLOTS] (fp64) += sum (act(z + (A_src G)/theta) - lambda/rho - gate)^2 for the
 *           thetas of `plan` (see admm_probe_plan) and, in slot 32, the same sum at G = 0, i.e. f(w);
 *           gates with done[g] != 0 are skipped (admm.py:316-325, 331-336).  scratch: 8*H*tc*ldn floats.
 *  select : per gate, replays `while f(beta) > est(beta, theta): theta *= 2` (admm.py:331-338) over the
 *           candidates of `plan` from the reduced sums (f(w) is taken from fk_acc, so both sides of the
 *           comparison come from the same kernel); writes theta_out[g] (already halved, admm.py:338) and
 *           done[g]; leaves the gate undecided if a lower bound failed or no candidate exits.  `done` has 12
 *           entries: [0,4) decided flags, [4,8) diagnostics of the last undecided pass (1 = bound not conclusive,
 *           2 = window exhausted, 3 = expansion not valid at k0), [8,12) the exponent concerned.  qmax: 4 floats
 *           (max |Q| per gate, zeroed before the pass, MAX-reduced over shards), used by moments plans only.
 *  apply  : w <- (0.5 rho T theta w - G)/(beta + 0.5 rho theta T)           (admm.py:340-343)
 * K = D for ADMM_SRC_X, H for ADMM_SRC_H. */
/* begin: once per `src` before the first admm_weight_grad of the phase (prepares the zstore refresh operands).
 * Tensor-core path: the x-phase gradient pass also measures max(1 + |lambda/rho| + |gate|) >= |R|, the bound that scales
 * the fp16 operand of the h-phase's A^T R GEMM -- run the x-phase of an iteration before its h-phase (admm.py:64-69 does). */
int admm_weight_begin(const admm_problem* p, int src, void* stream);
int admm_weight_grad(const admm_problem* p, int src, int t0, int tc, float* scratch,
                     double* g_acc, double* fw_acc, void* stream);
int admm_weight_finish_grad(const admm_problem* p, int src, const double* g_acc, float* grad_out,
                            double* est_acc, void* stream);
int admm_weight_probe(const admm_problem* p, int src, int t0, int tc, float* scratch, const float* grad,
                      const admm_probe_plan* plan, const int32_t* done, double* fk_acc, float* qmax, void* stream);
int admm_weight_select(const admm_problem* p, int src, const double* est_acc, const double* fk_acc, const float* qmax,
                       const admm_probe_plan* plan, int final_pass, int32_t* done, float* theta_out,
                       void* stream);
int admm_weight_apply(const admm_problem* p, int src, const float* grad, const float* theta,
                      void* stream);

/* admm.py:345-351 + 504-510 at one timestep: i,f,g,o,c,(h) then the dual ascent on i,f,g,o,c,
 * fused (a5,a6,a7,a9,a10 of SURVEY 8(a)).  For t == T, h_T / a / lambda_h are left to the three
 * admm_last_* calls below.  metrics: fp64 [ADMM_N_METRICS] accumulator or NULL. */
int admm_sweep_t(const admm_problem* p, int t, double* metrics, void* stream);

/* admm.py:439-487 / admm.no_dual_y.py:414-449 __update_primal_h at t = T:
 *  probe : sums[1 + 3*4] (fp64, zeroed) : S0 = ||h Wy - a||^2 and, for theta in {.1,.2,.4,.8},
 *          ||beta Wy - a||^2, <grad, beta - h>, ||beta - h||^2
 *  select: replays the while loop (admm.py:475-482) -> theta_out[0] in {.05,.1,.2,.4,.8}
 *  apply : h_T (admm.py:483-487), a (admm.py:489-502), lambda_h (admm.py:532-539),
 *          lambda_y if with_dual_y (admm.py:541-546); adds their metric terms. */
int admm_last_probe(const admm_problem* p, double* sums, void* stream);
int admm_last_select(const admm_problem* p, const double* sums, float* theta_out, void* stream);
int admm_last_apply(const admm_problem* p, const float* theta, double* metrics, void* stream);

/* Tensor-core (tcgen05 / TMA, 3xTF32) path for the gate GEMMs.  Returns 0 bytes when the shape is
 * not eligible (H % 64, ldn % 128, ...); otherwise the workspace the caller must provide in
 * admm_problem.tc_ws.  admm_tc_refresh() re-splits the weights after they change. */
#define ADMM_TC_WEIGHTS 1 /* wx, wh changed outside admm_weight_apply     */
#define ADMM_TC_INPUTS 2  /* x changed                                      */
#define ADMM_TC_STATE 4   /* h changed outside admm_forward_t / admm_sweep_t */
int64_t admm_tc_workspace_bytes(const admm_problem* p);
int admm_tc_refresh(const admm_problem* p, int what, void* stream);

/* Install new inputs for this shard from DEVICE buffers in the reference's layout (x_nm [n][T][D], y_nm [n][O], sample-major)
 * into the feature-major p->x / p->y, and refresh the tensor-core operand copies of x.  On the tensor-core path the kernel
 * also records, on the device, whether any value of x differs from what was installed before: if not, the stored
 * pre-activations (zstore) stay valid and the next x-phase gradient pass keeps its streaming form -- re-installing identical
 * inputs every step (an end-to-end loop that re-uploads its batch) costs one transposing copy, no GEMM pass and no host
 * round trip.  Leave admm_problem::z_valid as it is across this call. */
int admm_load_inputs(const admm_problem* p, const float* x_nm, const float* y_nm, void* stream);

/* Tensor-core path only: the fp16-pair operand of h (h 2^11 as hi + lo halves) represents |h| < 32.  h_t = (rho_h o tanh c -
 * lambda_h)/rho_h (admm.py:455-457) is below 1 in magnitude while o stays near a sigmoid value, but o is an unconstrained
 * ADMM primal: a larger value is clamped in the OPERAND (the fp32 state keeps it) and raises a sticky device flag.
 * Returns 1 if the flag is set, 0 if not (or no tensor-core workspace), negative on error; synchronises the stream;
 * reset != 0 clears it.  A set flag means the pre-activations computed since are wrong: use the CUDA-core path. */
int admm_tc_overflow(const admm_problem* p, int reset, void* stream);

/* Test hook: out[4][H][ldn] = pre-activations z_g = x_t W_g + h_{t-1} U_g at timestep t, through the
 * tensor-core path (use_tc != 0, needs tc_ws) or the CUDA-core path. */
int admm_debug_preact(const admm_problem* p, int t, float* out, int use_tc, void* stream);

/* Launch counter: number of kernels this library has launched since the last reset
 * (bench.py reports it as gpu_launches). */
int64_t admm_launch_count(int reset);

/* Per-kernel timing (measurement only; the reference has wall-clock timing of step() alone, demo.py:73-120,354-356).
 * admm_kernel_timing(1) makes every kernel launch of this library record a CUDA-event pair on its stream (returns the
 * previous setting; enabling drops earlier records); admm_kernel_timing_report() synchronises on the recorded events and
 * writes one line per kernel class, "name<TAB>launches<TAB>total_ms<LF>", NUL-terminated, into the HOST buffer buf[len],
 * clearing the records; returns the number of bytes written or a negative ADMM_E* code. */
int admm_kernel_timing(int enable);
int64_t admm_kernel_timing_report(char* host_buf, int64_t len);

/* ================================================================================================================
 * ADMM-LSTM-L: the linearised ADMM of
 * /root/reference/comparison_experiment/admm_l/admm_lstm.py driven by admm_l/main.py:139-191.
 *
 * Layout as above (feature-major, sample fastest).  The reference's per-timestep dictionaries z*[t], f[t] ... h[t],
 * t = 0..T-1, live in slot t+1 of [T+1][H][ldn] tensors; slot 0 is the all-zero h[-1] / c[-1] (main.py:87-88).
 * `base` carries x, y ([1][ldn]; the reference's W_y is [H,1], main.py:83), gate[] = i,f,g,o,c,h, the stacked weights
 * wx [4][D][H] / wh [4][H][H] in the order i,f,g,o, wy [H], a ([1][ldn]) and dual_y = lambda11 ([1][ldn]);
 * base.dual[] / base.dual_h / base.hp are unused.  Rows n >= base.n (ghost samples) are never touched and stay zero.
 * ================================================================================================================ */
typedef struct admm_l_hyper {
  float rho_s;    /* RHO_singular: rho1,3,5,7  (main.py:115)      */
  float rho_p;    /* RHO_plural:   rho2,4,6,8  (main.py:120)      */
  float rho9, rho10, rho11;                 /* main.py:125-129     */
  float lam_w, lam_u, lam_y;                /* lambda00, lambda02, lambda03 (main.py:112-114) */
  float n_norm;   /* the constant 4224 of update_a (admm_lstm.py:263-264) */
} admm_l_hyper;

typedef struct admm_l_problem {
  admm_problem base;
  float* z[4];      /* z_i z_f z_g z_o           : each [T+1][H][ldn] */
  float* lam_s[4];  /* lambda3, 1, 7, 5 (i f g o): z  = x W + h U     */
  float* lam_p[4];  /* lambda4, 2, 8, 6 (i f g o): gate = act(z)      */
  float* lam9;      /* c_t = f c_{t-1} + i g                          */
  float* lam10;     /* h_t = o tanh(c_t)                              */
  admm_l_hyper hp;
} admm_l_problem;

int admm_l_sizeof_problem(void);

/* main.py:85-103 at slot s = 1..T: z, gates, c, h (and the tensor-core side buffers of h); next_max[0..3] (zeroed by the
 * caller) receives max |gate_g| of the slot, the torch.max of update_z / update_zg in the first iteration.
 * scratch: 4*H*ldn floats. */
int admm_l_forward_t(const admm_l_problem* lp, int s, float* scratch, float* next_max, void* stream);
/* a = h_T W_y (main.py:102) */
int admm_l_output(const admm_l_problem* lp, void* stream);

/* Gram and right-hand-side sums of the weight subproblems (admm_lstm.py:107-163).  The residual
 * -z_t + x_t W + h_{t-1} U - lambda_t/rho is linear in W and U, so with V_g = z_g + lambda_g/rho
 *      acc_x[g] += sum_t x_t^T V_g,t  (g = 0..3, [D][H]),   acc_x[4] += sum_t x_t^T h_{t-1}   = S_xh
 *      acc_h[g] += sum_t h_{t-1}^T V_g,t      ([H][H]),      acc_h[4] += sum_t h_{t-1}^T h_{t-1} = S_hh
 * every gradient and every backtracking probe of the eight updates is O(K^2 H) algebra on these sums; they are what is
 * all-reduced across sample shards.  One call handles the timesteps [t0, t0+tc) (0-based, reference numbering);
 * scratch: 10*H*tc*ldn floats (V and h rows and their tf32 low parts).  gram_xx: sxx[D][D] += sum_t x_t^T x_t (constant
 * over the run).  sums_last: s_tt[H][H] += h_T^T h_T, p_t[H] += h_T^T (a + lambda11/rho11)  (update_Wy, :76-104). */
int admm_l_sums(const admm_l_problem* lp, int t0, int tc, float* scratch, double* acc_x, double* acc_h, void* stream);
int admm_l_gram_xx(const admm_l_problem* lp, double* sxx, void* stream);
int admm_l_sums_last(const admm_l_problem* lp, double* s_tt, double* p_t, void* stream);

/* The sweep at slot s (main.py:149-188): two elementwise kernels around the algorithm's mid-timestep reductions.
 *  gates : P = x W + h_{s-1} U for the four gates (gate GEMM -> scratch, 4*H*ldn floats); z_f,f, z_i,i, z_o,o, z_g,g in the
 *          reference's Gauss-Seidel order (:166-220) with red_max[g] = max |gate_g - lambda_p,g/rho_p| (update_z /
 *          update_zg, :168,:179; g = i,f,g,o) as INPUT; writes red_max[4] = max |(h - lambda10/rho10)/o| and
 *          red_sum[0] = sum o^2 (update_c, :225,:230)
 *  cell  : c (:223-241), h for s < T (:249-250) and the ten dual updates (:274-311); for s == T only c is written and
 *          admm_l_last() finishes the timestep.  next_max[0..3] receives max |gate_g - lambda_p,g/rho_p| of the values
 *          just written, i.e. the red_max[0..3] of slot s in the NEXT iteration (the reference's torch.max reads exactly
 *          these values before it updates them), so no separate pass over the state is needed.
 * red_max: 8 floats per slot (non-negative, combined with MAX across shards), red_sum: 1 double (SUM);
 * red_max[4..7], red_sum and next_max zeroed by the caller. */
int admm_l_sweep_gates(const admm_l_problem* lp, int s, float* scratch, float* red_max, double* red_sum, void* stream);
int admm_l_sweep_cell(const admm_l_problem* lp, int s, const float* scratch, const float* red_max, const double* red_sum,
                      float* next_max, void* stream);
/* s == T: h_T with theta_h[0] (update_h :251-258; the loop's exit is theta = smallest power of two >= max(1,
 * rho11 ||W_y||^2)), a (:262-266), lambda11 (:269-272), then the duals of slot T (next_max as above).
 * tmp: ldn floats; scratch still holds P of slot T. */
int admm_l_last(const admm_l_problem* lp, const float* theta_h, float* tmp, const float* scratch, float* next_max,
                void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADMM_LSTM_B200_H */
