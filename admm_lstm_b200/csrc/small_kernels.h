// small_kernels.h -- launchers of the non-GEMM kernels (see small_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include "../../include/admm_lstm_b200.h"

namespace admm {
int launch_wy_grad(const admm_problem& p, double* g_acc, cudaStream_t st);
int launch_wy_apply(const admm_problem& p, const double* g_acc, cudaStream_t st);
int launch_last_probe(const admm_problem& p, double* sums, cudaStream_t st);
int launch_last_select(const admm_problem& p, const double* sums, float* theta, cudaStream_t st);
int launch_last_apply(const admm_problem& p, const float* theta, double* metrics, cudaStream_t st);
int launch_output(const float* hT, const float* wy, float* a, int64_t ldn, int H, int O, cudaStream_t st);
int launch_weight_finish(const admm_problem& p, int src, const double* g_acc, float* grad, cudaStream_t st);
int launch_weight_est(const admm_problem& p, int src, const float* grad, double* est_acc, cudaStream_t st);
int launch_weight_select(const admm_problem& p, const double* est_acc, const double* fk_acc, const float* qmax,
                         const admm_probe_plan& plan, int final_pass, int32_t* done, float* theta, cudaStream_t st);
int launch_weight_apply(const admm_problem& p, int src, const float* grad, const float* theta, cudaStream_t st);
}  // namespace admm
