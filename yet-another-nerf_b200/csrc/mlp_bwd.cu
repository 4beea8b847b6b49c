// Backward of the fused NeRF-MLP (placeholder until the tcgen05 data-/weight-gradient kernels land).
#include <cuda_runtime.h>

#include "mlp_common.cuh"

extern "C" int64_t yn_mlp_bwd_workspace_bytes(const yn_mlp_arch* arch, int64_t n_points) {
  if (ynb::check_arch(arch) || n_points < 0) return -1;
  return 256;
}

extern "C" int yn_mlp_bwd(const yn_mlp_arch* arch, const float* directions, const float* rgb, const float* d_density,
                          const float* d_rgb, const float* params, const void* wpack, const float* aux,
                          const void* stash, void* workspace, float* grads, int64_t R, int P, void* stream) {
  return ynb::fail(YN_ERR_UNSUPPORTED, "yn_mlp_bwd: not implemented yet");
}
