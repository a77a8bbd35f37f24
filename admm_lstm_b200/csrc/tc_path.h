// tc_path.h -- tensor-core (tcgen05 / TMA / TMEM, 3xTF32) path of the gate GEMMs.
#pragma once
#include <cuda_runtime.h>

#include "../../include/admm_lstm_b200.h"
#include "gate_gemm.h"

namespace admm {
bool tc_eligible(const admm_problem* p);
int64_t tc_workspace_bytes(const admm_problem* p);
int tc_refresh_weights(const admm_problem* p, cudaStream_t st);
int tc_refresh_inputs(const admm_problem* p, cudaStream_t st);
int tc_refresh_state(const admm_problem* p, cudaStream_t st);
int tc_refresh_grad(const admm_problem* p, int src, const float* grad, cudaStream_t st);
int tc_refresh_wx_delta(const admm_problem* p, cudaStream_t st);
int gate_gemm_tc(int mode, const admm_problem* p, const GateGemmArgs& a, int tc, cudaStream_t st);
int atr_tc(const admm_problem* p, const AtrArgs& a, cudaStream_t st);
unsigned* tc_r_bound(const admm_problem* p);     // device slot: bound on |R| of the fp16 A^T R operand (bit pattern)
unsigned* tc_x_bound(const admm_problem* p);     // device slots [2]: the same bound for the next iteration, measured by the sweep
int tc_refresh_bound(const admm_problem* p, cudaStream_t st);      // ... recomputed from the state as it is
int tc_set_bound(const admm_problem* p, float v, cudaStream_t st);
int32_t* tc_z_dirty(const admm_problem* p);      // device flag: inputs changed since the z store was written
int tc_load_inputs(float* dst, const float* src, int64_t n, int64_t cols, int64_t ldn, int32_t* changed, cudaStream_t st);
unsigned* tc_h_overflow(const admm_problem* p);  // sticky device flag: an h value was clamped when its fp16 pair was written
void tc_h16(const admm_problem* p, __half** hi, __half** lo);   // fp16 pair of h 2^11, [T+1][H][ldn] each

// capi.cu helpers shared with admm_l.cu
int validate(const admm_problem* p, const char* who);
GateGemmArgs base_args(const admm_problem* p, int t_first);
int run_gate_gemm(int mode, const admm_problem* p, const GateGemmArgs& a, int tc, cudaStream_t st);
}  // namespace admm
