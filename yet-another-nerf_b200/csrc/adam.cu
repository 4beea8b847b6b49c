// torch.optim.Adam (scripts/run.py:159: lr from the config, betas (0.9, 0.999), eps 1e-8, no weight decay, no
// amsgrad) as ONE launch over the flat parameter / gradient / moment buffers of both networks.
#include <cuda_runtime.h>
#include <math.h>

#include "mlp_common.cuh"

namespace ynb {

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, int64_t n, float lr,
                                                   float b1, float b2, float eps, float bc1, float sqrt_bc2,
                                                   float gscale) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = g[i] * gscale;
    // exp_avg.lerp_(grad, 1 - beta1); exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    const float mi = m[i] + (gi - m[i]) * (1.f - b1);
    const float vi = v[i] * b2 + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / sqrt_bc2 + eps;
    p[i] = p[i] - (lr / bc1) * (mi / denom);
  }
}

}  // namespace ynb

// beta1 / beta2 are doubles: torch evaluates the bias corrections 1 - beta^step in double precision from the Python
// floats (0.999 as a float is 0.99900001: 1 - beta2 would be off by 1.3e-5 relative at step 1).
extern "C" int yn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                            float lr, double beta1, double beta2, float eps, int32_t step, float grad_scale,
                            void* stream) {
  if (n < 0 || step < 1) return ynb::fail(YN_ERR_INVALID_ARGUMENT, "yn_adam_step: n < 0 or step < 1");
  if (n == 0) return YN_OK;
  if (!params || !grads || !exp_avg || !exp_avg_sq) return ynb::fail(YN_ERR_INVALID_ARGUMENT, "yn_adam_step: null pointer");
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  ynb::adam_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      params, grads, exp_avg, exp_avg_sq, n, lr, (float)beta1, (float)beta2, eps, (float)bc1, (float)sqrt(bc2), grad_scale);
  return ynb::check_launch("yn_adam_step");
}

// Graph-friendly variant: the step counter and the learning rate live in device memory (state[0] = step as float,
// bumped by the caller on the stream; state[1] = lr), so a captured CUDA graph replays with fresh values.
namespace ynb {
__global__ void __launch_bounds__(256) adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                       float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                       const float* __restrict__ state, double beta1, double beta2,
                                                       float eps, float gscale) {
  // bias corrections in double like torch (one thread per block; fp32 `1 - powf(0.999f, 1)` is off by 1.3e-5 relative)
  __shared__ float s_bc[2];
  if (threadIdx.x == 0) {
    const double step = (double)state[0];
    s_bc[0] = (float)(1.0 - pow(beta1, step));
    s_bc[1] = (float)sqrt(1.0 - pow(beta2, step));
  }
  __syncthreads();
  const float lr = state[1], bc1 = s_bc[0], sqrt_bc2 = s_bc[1];
  const float b1 = (float)beta1, b2 = (float)beta2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = g[i] * gscale;
    const float mi = m[i] + (gi - m[i]) * (1.f - b1);
    const float vi = v[i] * b2 + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / sqrt_bc2 + eps;
    p[i] = p[i] - (lr / bc1) * (mi / denom);
  }
}
}  // namespace ynb

extern "C" int yn_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                                const float* state, double beta1, double beta2, float eps, float grad_scale, void* stream) {
  if (n < 0) return ynb::fail(YN_ERR_INVALID_ARGUMENT, "yn_adam_step_dev: n < 0");
  if (n == 0) return YN_OK;
  if (!params || !grads || !exp_avg || !exp_avg_sq || !state)
    return ynb::fail(YN_ERR_INVALID_ARGUMENT, "yn_adam_step_dev: null pointer");
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  ynb::adam_dev_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(params, grads, exp_avg, exp_avg_sq, n,
                                                                                      state, beta1, beta2, eps, grad_scale);
  return ynb::check_launch("yn_adam_step_dev");
}
