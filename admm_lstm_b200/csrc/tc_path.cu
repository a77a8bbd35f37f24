// tc_path.cu -- placeholder: the tcgen05 path is not wired yet; every shape reports "not eligible"
// so the fp32 CUDA-core kernels run.
#include "common.cuh"
#include "tc_path.h"

namespace admm {
bool tc_eligible(const admm_problem*) { return false; }
int64_t tc_workspace_bytes(const admm_problem*) { return 0; }
int tc_refresh_weights(const admm_problem*, cudaStream_t) { return ADMM_OK; }
int tc_refresh_grad(const admm_problem*, int, const float*, cudaStream_t) { return ADMM_OK; }
int gate_gemm_tc(int, const admm_problem*, const GateGemmArgs&, int, cudaStream_t) {
  set_error("tensor-core path not available");
  return ADMM_EINVAL;
}
int atr_tc(const admm_problem*, const AtrArgs&, cudaStream_t) {
  set_error("tensor-core path not available");
  return ADMM_EINVAL;
}
}  // namespace admm
