// capi.cu -- the extern "C" surface declared in include/admm_lstm_b200.h.
// Validates arguments, picks the kernel path (tcgen05 tensor-core path when a workspace is
// attached and the shape is eligible, fp32 CUDA-core path otherwise) and launches.  There is no
// CPU fallback: without an sm_100 device every compute entry point returns ADMM_ENODEV.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "gate_gemm.h"
#include "small_kernels.h"
#include "tc_path.h"

namespace admm {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
// ---- per-kernel CUDA-event timing (KernelScope, common.cuh) ------------------------------------------------------
namespace {
struct KtRec { const char* name; cudaEvent_t e0, e1; };
std::atomic<int> g_kt_on{0};
std::mutex g_kt_mu;
std::vector<KtRec> g_kt;
thread_local int g_kt_depth = 0;
}  // namespace

KernelScope::KernelScope(const char* name, cudaStream_t st) : name_(name), st_(st), e0_(nullptr) {
  if (!g_kt_on.load(std::memory_order_relaxed)) return;
  if (g_kt_depth++ > 0) return;                     // nested scopes: the outermost one owns the interval
  if (cudaEventCreate(&e0_) != cudaSuccess) { e0_ = nullptr; return; }
  cudaEventRecord(e0_, st_);
}
KernelScope::~KernelScope() {
  if (!g_kt_on.load(std::memory_order_relaxed) && !e0_) return;
  if (g_kt_depth > 0) --g_kt_depth;
  if (!e0_) return;
  cudaEvent_t e1;
  if (cudaEventCreate(&e1) != cudaSuccess) { cudaEventDestroy(e0_); return; }
  cudaEventRecord(e1, st_);
  std::lock_guard<std::mutex> lk(g_kt_mu);
  g_kt.push_back(KtRec{name_, e0_, e1});
}

int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return ADMM_ECUDA;
  }
  return ADMM_OK;
}

static int device_ok() {
  static int cached = -1;
  if (cached >= 0) return cached;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return cached = 0;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return cached = 0;
  return cached = (prop.major == 10) ? n : 0;
}

int validate(const admm_problem* p, const char* who) {
  if (!p) { set_error("%s: null problem", who); return ADMM_EINVAL; }
  if (!device_ok()) { set_error("%s: no sm_100 CUDA device (this library has no CPU path)", who); return ADMM_ENODEV; }
  ADMM_REQUIRE(p->n > 0 && p->n_global >= p->n, "%s: bad n=%lld n_global=%lld", who, (long long)p->n, (long long)p->n_global);
  ADMM_REQUIRE(p->ldn >= p->n && p->ldn % 128 == 0, "%s: ldn=%lld must be a multiple of 128 and >= n", who, (long long)p->ldn);
  ADMM_REQUIRE(p->T > 0 && p->D > 0 && p->H > 0 && p->O > 0, "%s: bad sizes T=%d D=%d H=%d O=%d", who, p->T, p->D, p->H, p->O);
  ADMM_REQUIRE(p->O <= ADMM_MAX_O, "%s: output_size %d > %d", who, p->O, ADMM_MAX_O);
  ADMM_REQUIRE(p->variant == ADMM_VARIANT_ADMM || p->variant == ADMM_VARIANT_NO_DUAL_Y, "%s: bad variant", who);
  ADMM_REQUIRE(!(p->with_dual_y && p->variant == ADMM_VARIANT_NO_DUAL_Y), "%s: with_dual_y needs the admm variant", who);
  return ADMM_OK;
}

GateGemmArgs base_args(const admm_problem* p, int t_first) {
  // t_first = first timestep t (1-based) of the launch
  GateGemmArgs a;
  memset(&a, 0, sizeof(a));
  const int64_t slab = (int64_t)p->H * p->ldn;
  a.n = p->n; a.ldn = p->ldn; a.D = p->D; a.H = p->H;
  a.x = p->x + (int64_t)(t_first - 1) * p->D * p->ldn;
  a.h_prev = p->gate[5] + (int64_t)(t_first - 1) * slab;
  a.c_prev = p->gate[4] + (int64_t)(t_first - 1) * slab;
  a.x_tstride = (int64_t)p->D * p->ldn;
  a.s_tstride = slab;
  a.wx = p->wx; a.wh = p->wh;
  for (int q = 0; q < 6; ++q) a.gate[q] = p->gate[q] + (int64_t)t_first * slab;
  for (int q = 0; q < 5; ++q) a.dual[q] = p->dual[q] + (int64_t)t_first * slab;
  a.dual_h = p->dual_h;
  a.rho = make_rho(p->hp);
  a.tc = 1;
  a.tc_ws = p->tc_ws;
  return a;
}

int run_gate_gemm(int mode, const admm_problem* p, const GateGemmArgs& a, int tc, cudaStream_t st) {
  if (p->tc_ws && tc_eligible(p)) return gate_gemm_tc(mode, p, a, tc, st);
  return gate_gemm_simt(mode, a, tc, st);
}

}  // namespace admm

using namespace admm;

extern "C" {

const char* admm_last_error(void) { return g_err; }
int admm_abi_version(void) { return ADMM_ABI_VERSION; }
int admm_sizeof_problem(void) { return (int)sizeof(admm_problem); }
int admm_device_ok(void) { return device_ok(); }
int64_t admm_launch_count(int reset) {
  return reset ? g_launches.exchange(0) : g_launches.load();
}

int admm_kernel_timing(int enable) {
  const int was = g_kt_on.exchange(enable ? 1 : 0);
  if (enable && !was) {
    std::lock_guard<std::mutex> lk(g_kt_mu);
    for (auto& r : g_kt) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    g_kt.clear();
  }
  return was;
}

int64_t admm_kernel_timing_report(char* buf, int64_t len) {
  if (!buf || len <= 0) return ADMM_EINVAL;
  std::map<std::string, std::pair<int64_t, double>> agg;
  std::vector<std::string> order;
  {
    std::lock_guard<std::mutex> lk(g_kt_mu);
    for (auto& r : g_kt) {
      if (cudaEventSynchronize(r.e1) != cudaSuccess) { cudaGetLastError(); continue; }
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) { cudaGetLastError(); continue; }
      auto it = agg.find(r.name);
      if (it == agg.end()) { order.push_back(r.name); it = agg.emplace(r.name, std::make_pair((int64_t)0, 0.0)).first; }
      it->second.first += 1;
      it->second.second += ms;
      cudaEventDestroy(r.e0);
      cudaEventDestroy(r.e1);
    }
    g_kt.clear();
  }
  std::string out;
  char line[256];
  for (auto& name : order) {
    snprintf(line, sizeof(line), "%s\t%lld\t%.6f\n", name.c_str(), (long long)agg[name].first, agg[name].second);
    out += line;
  }
  if ((int64_t)out.size() + 1 > len) { set_error("admm_kernel_timing_report: buffer of %lld bytes too small", (long long)len); return ADMM_EINVAL; }
  memcpy(buf, out.c_str(), out.size() + 1);
  return (int64_t)out.size();
}

int admm_forward_t(const admm_problem* p, int t, void* stream) {
  int rc = validate(p, "admm_forward_t");
  if (rc) return rc;
  ADMM_REQUIRE(t >= 1 && t <= p->T, "admm_forward_t: t=%d out of 1..%d", t, p->T);
  cudaStream_t st = (cudaStream_t)stream;
  GateGemmArgs a = base_args(p, t);
  if (p->tc_ws && tc_eligible(p) && p->zstore) { a.zstore = p->zstore; a.zT = p->T; a.zt0 = t - 1; }
  rc = run_gate_gemm(GG_FORWARD, p, a, 1, st);
  if (rc) return rc;
  // the forward pass is the initialisation (admm.py:164-173: duals are zero, gates are activations): |R| <= 2
  if (t == p->T && p->tc_ws && tc_eligible(p) && (rc = tc_set_bound(p, 2.0f, st))) return rc;
  if (t == p->T && p->a)
    return launch_output(p->gate[5] + (int64_t)p->T * p->H * p->ldn, p->wy, p->a, p->ldn, p->H, p->O, st);
  return ADMM_OK;
}

int admm_predict(const admm_problem* p, float* work, float* out, void* stream) {
  int rc = validate(p, "admm_predict");
  if (rc) return rc;
  ADMM_REQUIRE(work && out, "admm_predict: null buffers");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t slab = (int64_t)p->H * p->ldn;
  float* hbuf[2] = {work, work + slab};
  float* cbuf[2] = {work + 2 * slab, work + 3 * slab};
  if (cudaMemsetAsync(work, 0, sizeof(float) * 4 * slab, st) != cudaSuccess) return check_launch("predict memset");
  for (int t = 1; t <= p->T; ++t) {
    GateGemmArgs a;
    memset(&a, 0, sizeof(a));
    a.n = p->n; a.ldn = p->ldn; a.D = p->D; a.H = p->H;
    a.x = p->x + (int64_t)(t - 1) * p->D * p->ldn;
    a.h_prev = hbuf[(t - 1) & 1];
    a.c_prev = cbuf[(t - 1) & 1];
    a.wx = p->wx; a.wh = p->wh;
    a.gate[4] = cbuf[t & 1];
    a.gate[5] = hbuf[t & 1];
    a.rho = make_rho(p->hp);
    a.tc = 1;
    rc = gate_gemm_simt(GG_FORWARD, a, 1, st);
    if (rc) return rc;
  }
  return launch_output(hbuf[p->T & 1], p->wy, out, p->ldn, p->H, p->O, st);
}

int admm_wy_grad(const admm_problem* p, double* g_acc, void* stream) {
  int rc = validate(p, "admm_wy_grad");
  if (rc) return rc;
  ADMM_REQUIRE(g_acc, "admm_wy_grad: null accumulator");
  return launch_wy_grad(*p, g_acc, (cudaStream_t)stream);
}

int admm_wy_apply(const admm_problem* p, const double* g_acc, void* stream) {
  int rc = validate(p, "admm_wy_apply");
  if (rc) return rc;
  return launch_wy_apply(*p, g_acc, (cudaStream_t)stream);
}

int admm_weight_begin(const admm_problem* p, int src, void* stream) {
  int rc = validate(p, "admm_weight_begin");
  if (rc) return rc;
  ADMM_REQUIRE(src == ADMM_SRC_X || src == ADMM_SRC_H, "admm_weight_begin: bad src");
  const bool use_tc = p->tc_ws && tc_eligible(p);
  if (src == ADMM_SRC_X && use_tc) {
    // the bound on |R| that scales the fp16 operand of both A^T R GEMMs: measured by the previous sweep over the lambda / gate
    // values it wrote (or by the forward initialisation / admm_tc_refresh(ADMM_TC_STATE)); they do not change in the weight phase
    if (cudaMemcpyAsync(tc_r_bound(p), tc_x_bound(p), sizeof(unsigned), cudaMemcpyDeviceToDevice, (cudaStream_t)stream) !=
        cudaSuccess)
      return check_launch("r_bound copy");
  }
  if (src == ADMM_SRC_H && p->zstore && p->wx_prev && use_tc) return tc_refresh_wx_delta(p, (cudaStream_t)stream);
  return ADMM_OK;
}

int admm_weight_grad(const admm_problem* p, int src, int t0, int tc, float* scratch, double* g_acc,
                     double* fw_acc, void* stream) {
  int rc = validate(p, "admm_weight_grad");
  if (rc) return rc;
  ADMM_REQUIRE(src == ADMM_SRC_X || src == ADMM_SRC_H, "admm_weight_grad: bad src");
  ADMM_REQUIRE(t0 >= 0 && tc >= 1 && t0 + tc <= p->T, "admm_weight_grad: bad timestep range %d+%d", t0, tc);
  ADMM_REQUIRE(scratch && g_acc && fw_acc, "admm_weight_grad: null buffers");
  cudaStream_t st = (cudaStream_t)stream;
  GateGemmArgs a = base_args(p, t0 + 1);
  const bool use_tc = p->tc_ws && tc_eligible(p);
  const bool atr_on_tc = use_tc;
  a.scratch = scratch; a.tc = tc; a.fw_acc = fw_acc; a.src = src;
  a.scratch_q = atr_on_tc ? scratch + 4LL * p->H * tc * p->ldn : nullptr;     // tf32 low part of R^T
  // tensor-core path: R^T is written as fp16 pairs scaled from the bound 1 + max(|lambda/rho| + |gate|) >= |R| (known before
  // the phase starts, admm_weight_begin) and the fp16 A^T R GEMM runs; an x-phase with fewer than 8 input features keeps
  // fp32 R for the CUDA-core reduction
  const bool r16 = use_tc && (src == ADMM_SRC_H || p->D >= 8);
  if (r16) {
    a.r_bound = tc_r_bound(p);
    a.r16_hi = reinterpret_cast<__half*>(scratch);
    a.r16_lo = reinterpret_cast<__half*>(a.scratch_q);
  }
  if (use_tc && p->zstore && p->wx_prev) {
    // x-phase: full GEMM, z kept; h-phase: z <- z + x (W_new - W_old), no full GEMM
    a.zstore = p->zstore; a.zT = p->T; a.zt0 = t0;
    a.z_accumulate = (src == ADMM_SRC_H);
  }
  if (use_tc && p->zstore && p->wx_prev && src == ADMM_SRC_X && p->z_valid) {
    // the previous sweep (or the forward initialisation) left z of the current weights and states in zstore:
    // no GEMM, one streaming pass
    GradFromZArgs e;
    e.n = p->n; e.ldn = p->ldn; e.H = p->H; e.tc = tc; e.zT = p->T; e.zt0 = t0; e.zstore = p->zstore;
    for (int g = 0; g < 4; ++g) { e.gate[g] = a.gate[g]; e.dual[g] = a.dual[g]; e.rho[g] = p->hp.rho[g]; }
    e.s_tstride = a.s_tstride;
    e.r = a.scratch; e.r_lo = a.scratch_q; e.fw_acc = fw_acc; e.bound_track = nullptr;
    e.r16_hi = a.r16_hi; e.r16_lo = a.r16_lo; e.r_bound = a.r_bound;
    // exactly one of the two passes runs, decided on the device: the streaming pass over the stored z unless the inputs were
    // replaced by different values since it was written (admm_load_inputs), else the pass with its own GEMM (it rewrites z)
    e.skip_if = tc_z_dirty(p);
    rc = grad_from_z(e, st);
    if (rc) return rc;
    a.run_if = tc_z_dirty(p);
    rc = run_gate_gemm(GG_GRAD, p, a, tc, st);
  } else {
    rc = run_gate_gemm(GG_GRAD, p, a, tc, st);
  }
  if (rc) return rc;
  AtrArgs r;
  memset(&r, 0, sizeof(r));
  r.ldn = p->ldn; r.H = p->H; r.tc = tc;
  r.scratch_lo = a.scratch_q;
  r.K = (src == ADMM_SRC_X) ? p->D : p->H;
  r.a_src = (src == ADMM_SRC_X) ? a.x : a.h_prev;
  r.a_tstride = (int64_t)r.K * p->ldn;
  r.scratch = scratch; r.g_acc = g_acc;
  if (r16) { r.r16_hi = a.r16_hi; r.r16_lo = a.r16_lo; r.r_bound = a.r_bound; }
  if (atr_on_tc) return atr_tc(p, r, st);
  return atr_simt(r, st);
}

int admm_weight_finish_grad(const admm_problem* p, int src, const double* g_acc, float* grad_out, double* est_acc,
                            void* stream) {
  int rc = validate(p, "admm_weight_finish_grad");
  if (rc) return rc;
  ADMM_REQUIRE(g_acc && grad_out && est_acc, "admm_weight_finish_grad: null buffers");
  rc = launch_weight_finish(*p, src, g_acc, grad_out, (cudaStream_t)stream);
  if (rc) return rc;
  rc = launch_weight_est(*p, src, grad_out, est_acc, (cudaStream_t)stream);
  if (rc) return rc;
  if (p->tc_ws && tc_eligible(p)) return tc_refresh_grad(p, src, grad_out, (cudaStream_t)stream);
  return ADMM_OK;
}

static int check_plan(const admm_probe_plan* plan, const char* who) {
  ADMM_REQUIRE(plan, "%s: null plan", who);
  ADMM_REQUIRE(plan->ncand >= (plan->moments ? 0 : 1) && plan->ncand <= ADMM_MAX_CAND, "%s: bad ncand %d", who, plan->ncand);
  for (int g = 0; g < 4; ++g) {
    ADMM_REQUIRE(plan->k0[g] >= 0 && plan->k0[g] + (plan->moments ? 0 : plan->ncand) <= ADMM_EST_CAND, "%s: bad k0[%d]=%d", who,
                 g, plan->k0[g]);
    ADMM_REQUIRE(!plan->proof || plan->k0[g] + (plan->moments ? plan->ncand : 0) <= ADMM_MAX_CAND, "%s: proof range too long",
                 who);
  }
  return ADMM_OK;
}

int admm_weight_probe(const admm_problem* p, int src, int t0, int tc, float* scratch, const float* grad,
                      const admm_probe_plan* plan, const int32_t* done, double* fk_acc, float* qmax, void* stream) {
  int rc = validate(p, "admm_weight_probe");
  if (rc) return rc;
  ADMM_REQUIRE(src == ADMM_SRC_X || src == ADMM_SRC_H, "admm_weight_probe: bad src");
  ADMM_REQUIRE(t0 >= 0 && tc >= 1 && t0 + tc <= p->T, "admm_weight_probe: bad timestep range");
  if ((rc = check_plan(plan, "admm_weight_probe"))) return rc;
  ADMM_REQUIRE(scratch && grad && done && fk_acc, "admm_weight_probe: null buffers");
  ADMM_REQUIRE(!plan->moments || qmax, "admm_weight_probe: a moments plan needs qmax");
  cudaStream_t st = (cudaStream_t)stream;
  GateGemmArgs a = base_args(p, t0 + 1);
  const int64_t half = 4LL * p->H * tc * p->ldn;
  a.tc = tc; a.src = src; a.grad = grad; a.done = done;
  a.scratch = scratch; a.scratch_q = scratch + half;
  const bool stored = p->tc_ws && tc_eligible(p) && p->zstore && p->wx_prev;
  if (stored) { a.zstore = p->zstore; a.zT = p->T; a.zt0 = t0; }
  if (stored && plan->moments) {
    // fused moment pass: Q = A_src G stays in TMEM, the GEMM's epilogue accumulates the moment sums and (on one unit per
    // tile) the lower-bound proofs below the expansion -- no Q round trip through HBM, no separate proof passes
    static const int fused = [] {
      const char* e = getenv("ADMM_FUSED_MOMENTS");      // A/B switch for measurements
      return e ? atoi(e) : 1;
    }();
    bool ok = fused != 0;
    for (int g = 0; g < 4; ++g) {
      a.mom_k0[g] = plan->k0[g];
      a.mom_order = (plan->order == 4) ? 4 : 6;
      a.mom_pc[g] = plan->proof ? plan->k0[g] + plan->ncand : 0;
      ok = ok && a.mom_pc[g] <= 16 && plan->k0[g] < 64;
    }
    if (ok) {
      a.fk_acc = fk_acc; a.qmax = qmax;
      return run_gate_gemm(GG_MOMENTS, p, a, tc, st);
    }
  }
  rc = run_gate_gemm(GG_PROBE, p, a, tc, st);
  if (rc) return rc;
  ProbeEvalArgs e;
  e.n = p->n; e.ldn = p->ldn; e.H = p->H; e.tc = tc;
  e.z0 = stored ? p->zstore : a.scratch; e.q = a.scratch_q;
  e.z_T = stored ? p->T : tc; e.z_t0 = stored ? t0 : 0;
  for (int g = 0; g < 4; ++g) { e.gate[g] = a.gate[g]; e.dual[g] = a.dual[g]; e.rho[g] = p->hp.rho[g]; }
  e.s_tstride = a.s_tstride;
  e.done = done; e.fk_acc = fk_acc;
  // window: all units, theta = 2^(k0[g] + c)
  for (int g = 0; g < 4; ++g) { e.kbase[g] = plan->k0[g]; e.nc[g] = plan->ncand; e.slot0[g] = 0; }
  e.jmod = 1; e.jrem = 0; e.publish_fw = 1;
  rc = plan->moments ? probe_moments(e, qmax, st) : probe_eval(e, st);
  if (rc || !plan->proof) return rc;
  if (plan->moments) {
    // insurance below the expansion: lower bounds of f for k < k0 + ncand on 1/8 of the units (these exponents are many
    // doublings below the exit, so the bound exceeds est by orders of magnitude; see weight_select_kernel)
    for (int g = 0; g < 4; ++g) { e.kbase[g] = 0; e.nc[g] = plan->k0[g] + plan->ncand; e.slot0[g] = ADMM_MAX_CAND + 1; }
    e.jmod = 8; e.jrem = 0; e.publish_fw = 0;
    return probe_eval(e, st);
  }
  // lower bounds below the window: every k < k0[g] on unit blocks 0 mod 8, and the three exponents next to the
  // window additionally on the odd blocks (disjoint sets, so the sums add up to one partial sum over 5/8 of the units)
  for (int g = 0; g < 4; ++g) { e.kbase[g] = 0; e.nc[g] = plan->k0[g]; e.slot0[g] = ADMM_MAX_CAND + 1; }
  e.jmod = 8; e.jrem = 0; e.publish_fw = 0;
  rc = probe_eval(e, st);
  if (rc) return rc;
  for (int g = 0; g < 4; ++g) {
    e.kbase[g] = plan->k0[g] > 3 ? plan->k0[g] - 3 : 0;
    e.nc[g] = plan->k0[g] - e.kbase[g];
    e.slot0[g] = ADMM_MAX_CAND + 1 + e.kbase[g];
  }
  e.jmod = 2; e.jrem = 1;
  return probe_eval(e, st);
}

int admm_weight_select(const admm_problem* p, int src, const double* est_acc, const double* fk_acc, const float* qmax,
                       const admm_probe_plan* plan, int final_pass, int32_t* done, float* theta_out, void* stream) {
  int rc = validate(p, "admm_weight_select");
  if (rc) return rc;
  (void)src;
  if ((rc = check_plan(plan, "admm_weight_select"))) return rc;
  ADMM_REQUIRE(est_acc && fk_acc && done && theta_out, "admm_weight_select: null buffers");
  ADMM_REQUIRE(!plan->moments || qmax, "admm_weight_select: a moments plan needs qmax");
  return launch_weight_select(*p, est_acc, fk_acc, qmax, *plan, final_pass, done, theta_out, (cudaStream_t)stream);
}

int admm_weight_apply(const admm_problem* p, int src, const float* grad, const float* theta, void* stream) {
  int rc = validate(p, "admm_weight_apply");
  if (rc) return rc;
  if (src == ADMM_SRC_X && p->wx_prev) {
    if (cudaMemcpyAsync(p->wx_prev, p->wx, sizeof(float) * 4 * p->D * p->H, cudaMemcpyDeviceToDevice,
                        (cudaStream_t)stream) != cudaSuccess)
      return check_launch("wx_prev copy");
  }
  rc = launch_weight_apply(*p, src, grad, theta, (cudaStream_t)stream);
  if (rc) return rc;
  if (p->tc_ws && tc_eligible(p)) return tc_refresh_weights(p, (cudaStream_t)stream);
  return ADMM_OK;
}

int admm_sweep_t(const admm_problem* p, int t, double* metrics, void* stream) {
  int rc = validate(p, "admm_sweep_t");
  if (rc) return rc;
  ADMM_REQUIRE(t >= 1 && t <= p->T, "admm_sweep_t: t=%d out of 1..%d", t, p->T);
  GateGemmArgs a = base_args(p, t);
  a.last = (t == p->T);
  a.metrics = metrics;
  // keep z_t: it is the pre-activation the next iteration's x-phase gradient starts from (admm_problem::z_valid)
  if (p->tc_ws && tc_eligible(p) && p->zstore) { a.zstore = p->zstore; a.zT = p->T; a.zt0 = t - 1; }
  const bool use_tc = p->tc_ws && tc_eligible(p);
  cudaStream_t st = (cudaStream_t)stream;
  if (use_tc) {
    // the sweep measures the bound on |R| of the next iteration's gradient passes while it writes lambda and the gates
    a.xbound_track = tc_x_bound(p);
    if (t == 1 && cudaMemsetAsync(a.xbound_track + 1, 0, sizeof(unsigned), st) != cudaSuccess) return check_launch("x_bound memset");
  }
  rc = run_gate_gemm(GG_SWEEP, p, a, 1, st);
  if (rc || !use_tc || t != p->T) return rc;
  if (cudaMemcpyAsync(a.xbound_track, a.xbound_track + 1, sizeof(unsigned), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
    return check_launch("x_bound publish");
  // the sweep rewrote every stored pre-activation from the current inputs
  if (cudaMemsetAsync(tc_z_dirty(p), 0, sizeof(int32_t), st) != cudaSuccess) return check_launch("z_dirty clear");
  return ADMM_OK;
}

int admm_last_probe(const admm_problem* p, double* sums, void* stream) {
  int rc = validate(p, "admm_last_probe");
  if (rc) return rc;
  return launch_last_probe(*p, sums, (cudaStream_t)stream);
}
int admm_last_select(const admm_problem* p, const double* sums, float* theta_out, void* stream) {
  int rc = validate(p, "admm_last_select");
  if (rc) return rc;
  return launch_last_select(*p, sums, theta_out, (cudaStream_t)stream);
}
int admm_last_apply(const admm_problem* p, const float* theta, double* metrics, void* stream) {
  int rc = validate(p, "admm_last_apply");
  if (rc) return rc;
  return launch_last_apply(*p, theta, metrics, (cudaStream_t)stream);
}

int64_t admm_tc_workspace_bytes(const admm_problem* p) {
  if (!p) return 0;
  return tc_workspace_bytes(p);
}
int admm_tc_refresh(const admm_problem* p, int what, void* stream) {
  int rc = validate(p, "admm_tc_refresh");
  if (rc) return rc;
  if (!(p->tc_ws && tc_eligible(p))) return ADMM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // inputs first: the operand scale of x enters the weights' scales (tc_refresh_inputs re-prepares the weights itself)
  if ((what & ADMM_TC_INPUTS) && (rc = tc_refresh_inputs(p, st))) return rc;
  if ((what & ADMM_TC_WEIGHTS) && !(what & ADMM_TC_INPUTS) && (rc = tc_refresh_weights(p, st))) return rc;
  if ((what & ADMM_TC_STATE) && (rc = tc_refresh_state(p, st))) return rc;
  if ((what & ADMM_TC_STATE) && (rc = tc_refresh_bound(p, st))) return rc;
  return ADMM_OK;
}

int admm_load_inputs(const admm_problem* p, const float* x_nm, const float* y_nm, void* stream) {
  int rc = validate(p, "admm_load_inputs");
  if (rc) return rc;
  ADMM_REQUIRE(x_nm && y_nm, "admm_load_inputs: null inputs");
  cudaStream_t st = (cudaStream_t)stream;
  const bool use_tc = p->tc_ws && tc_eligible(p);
  int32_t* changed = use_tc ? tc_z_dirty(p) : nullptr;
  rc = tc_load_inputs(const_cast<float*>(p->x), x_nm, p->n, (int64_t)p->T * p->D, p->ldn, changed, st);
  if (rc) return rc;
  rc = tc_load_inputs(const_cast<float*>(p->y), y_nm, p->n, p->O, p->ldn, nullptr, st);
  if (rc || !use_tc) return rc;
  return tc_refresh_inputs(p, st);          // operand copies of x (and the weight scales tied to its scale)
}

int admm_tc_overflow(const admm_problem* p, int reset, void* stream) {
  int rc = validate(p, "admm_tc_overflow");
  if (rc) return rc;
  if (!(p->tc_ws && tc_eligible(p))) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned flag = 0;
  if (cudaMemcpyAsync(&flag, tc_h_overflow(p), sizeof(flag), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess)
    return check_launch("admm_tc_overflow") ? ADMM_ECUDA : ADMM_ECUDA;
  if (reset && flag && cudaMemsetAsync(tc_h_overflow(p), 0, sizeof(flag), st) != cudaSuccess) return check_launch("overflow reset");
  return flag ? 1 : 0;
}

int admm_debug_preact(const admm_problem* p, int t, float* out, int use_tc, void* stream) {
  int rc = validate(p, "admm_debug_preact");
  if (rc) return rc;
  ADMM_REQUIRE(t >= 1 && t <= p->T && out, "admm_debug_preact: bad arguments");
  GateGemmArgs a = base_args(p, t);
  a.scratch = out; a.tc = 1;
  if (use_tc) {
    ADMM_REQUIRE(p->tc_ws && tc_eligible(p), "admm_debug_preact: tensor-core path not available for this problem");
    return gate_gemm_tc(GG_RAWZ, p, a, 1, (cudaStream_t)stream);
  }
  return gate_gemm_simt(GG_RAWZ, a, 1, (cudaStream_t)stream);
}

}  // extern "C"
