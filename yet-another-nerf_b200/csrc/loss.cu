// Per-image rgb losses of the training step in one launch each way: ground-truth gather at the sampled pixels
// (`sample_grid`, yanerf/pipelines/utils.py:272-296: flat index = x + W_img * y) fused with the squared-error mean of
// `_rgb_metrics` (pipelines/utils.py:137-158): mse_b = mean over (rays x channels) of (pred - gt)^2, shape (B,), and
// huber_b = (sqrt(max(1 + mse_b / 0.03^2, 0) + 1e-4) - 1) * 0.03 (189-203).  Fixed-order two-level reduction: the result
// is deterministic.  The backward writes d(pred) for incoming d(mse_b), d(huber_b).
#include <cuda_runtime.h>

#include "mlp_common.cuh"

namespace ynb {


struct LossParams {
  const float* pred;   // [B, n, C]
  const float* image;  // [B, H, W, C]
  const float* xy;     // [B, n, 2] float pixel coordinates
  float* mse;          // [B]
  float* huber;        // [B]
  const float* g_mse;  // backward: [B] (may be null)
  const float* g_huber;
  float* d_pred;       // [B, n, C]
  int64_t n;
  int C, width, height;
};

__device__ __forceinline__ int64_t pixel_index(const LossParams& p, int64_t b, int64_t i) {
  const float x = __ldg(p.xy + (b * p.n + i) * 2), y = __ldg(p.xy + (b * p.n + i) * 2 + 1);
  return (int64_t)__fadd_rn(x, __fmul_rn((float)p.width, y));  // (x + W * y).long()
}

// grid (nb, B): block (k, b) sums its strided share of image b's rays in double precision into partial[b][k]; the block
// that finishes LAST for image b (device counter) folds the nb partial sums in index order: the result does not depend on
// the order in which the blocks ran.  nb = 16 for training batches, 128 for full-image grids (a function of n alone).
constexpr int kLossBlocks = 128;  // row length of `partial`
__host__ __device__ inline int loss_blocks(int64_t n) { return n > 65536 ? kLossBlocks : 16; }
__global__ void __launch_bounds__(256) rgb_loss_fwd_kernel(const LossParams p, double* __restrict__ partial, unsigned int* __restrict__ counter) {
  __shared__ double s_part[8];
  __shared__ bool s_last;
  const int64_t b = blockIdx.y;
  const int64_t n_pix = (int64_t)p.width * p.height;
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = pixel_index(p, b, i);
    const float* gt = p.image + (b * n_pix + pix) * p.C;
    const float* pr = p.pred + (b * p.n + i) * p.C;
    for (int c = 0; c < p.C; ++c) {
      const float d = __ldg(pr + c) - __ldg(gt + c);
      acc += (double)(d * d);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < 8; ++w) tot += s_part[w];
    partial[b * kLossBlocks + blockIdx.x] = tot;
    __threadfence();
    s_last = atomicAdd(counter + b, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    double tot = 0.0;
    for (int k = 0; k < (int)gridDim.x; ++k) tot += reinterpret_cast<volatile double*>(partial)[b * kLossBlocks + k];
    const float mse = (float)(tot / (double)(p.n * p.C));
    p.mse[b] = mse;
    p.huber[b] = (sqrtf(fmaxf(1.f + mse / (0.03f * 0.03f), 0.f) + 1e-4f) - 1.f) * 0.03f;
    counter[b] = 0u;  // ready for the next launch (the scratch is caller-owned and zeroed once)
  }
}

__global__ void __launch_bounds__(256) rgb_loss_bwd_kernel(const LossParams p, int64_t B) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * p.n) return;
  const int64_t b = t / p.n, i = t % p.n;
  const int64_t n_pix = (int64_t)p.width * p.height;
  // d loss / d mse_b: direct + through huber = 0.03 * sqrt(max(1 + mse / s^2, 0) + eps) - 0.03
  float g = p.g_mse ? __ldg(p.g_mse + b) : 0.f;
  if (p.g_huber) {
    const float mse = __ldg(p.mse + b);
    const float inner = 1.f + mse / (0.03f * 0.03f);
    if (inner > 0.f) g += __ldg(p.g_huber + b) * 0.03f * 0.5f / sqrtf(inner + 1e-4f) / (0.03f * 0.03f);
  }
  const float k = 2.f * g / (float)(p.n * p.C);
  const int64_t pix = pixel_index(p, b, i);
  const float* gt = p.image + (b * n_pix + pix) * p.C;
  const float* pr = p.pred + t * p.C;
  for (int c = 0; c < p.C; ++c) p.d_pred[t * p.C + c] = k * (__ldg(pr + c) - __ldg(gt + c));
}

}  // namespace ynb

extern "C" int64_t yn_rgb_loss_scratch_bytes(int64_t B) { return B < 0 ? -1 : B * (ynb::kLossBlocks * 8 + 8); }

extern "C" int yn_rgb_loss_fwd(const float* pred, const float* image, const float* xy, float* mse, float* huber, void* scratch,
                               int64_t B, int64_t n, int C, int width, int height, void* stream) {
  if (B < 0 || n < 1 || C < 1 || width < 1 || height < 1) return ynb::fail(YN_ERR_INVALID_ARGUMENT, "yn_rgb_loss_fwd: bad sizes");
  if (B == 0) return YN_OK;
  if (!pred || !image || !xy || !mse || !huber || !scratch) return ynb::fail(YN_ERR_INVALID_ARGUMENT, "yn_rgb_loss_fwd: null pointer");
  ynb::LossParams p = {};
  p.pred = pred; p.image = image; p.xy = xy; p.mse = mse; p.huber = huber; p.n = n; p.C = C; p.width = width; p.height = height;
  double* partial = static_cast<double*>(scratch);
  unsigned int* counter = reinterpret_cast<unsigned int*>(static_cast<uint8_t*>(scratch) + B * ynb::kLossBlocks * 8);
  ynb::rgb_loss_fwd_kernel<<<dim3(ynb::loss_blocks(n), (unsigned)B), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, partial, counter);
  return ynb::check_launch("yn_rgb_loss_fwd");
}

extern "C" int yn_rgb_loss_bwd(const float* pred, const float* image, const float* xy, const float* mse, const float* g_mse,
                               const float* g_huber, float* d_pred, int64_t B, int64_t n, int C, int width, int height,
                               void* stream) {
  if (B < 0 || n < 1 || C < 1 || width < 1 || height < 1) return ynb::fail(YN_ERR_INVALID_ARGUMENT, "yn_rgb_loss_bwd: bad sizes");
  if (B == 0) return YN_OK;
  if (!pred || !image || !xy || !mse || !d_pred) return ynb::fail(YN_ERR_INVALID_ARGUMENT, "yn_rgb_loss_bwd: null pointer");
  ynb::LossParams p = {};
  p.pred = pred; p.image = image; p.xy = xy; p.mse = const_cast<float*>(mse); p.g_mse = g_mse; p.g_huber = g_huber; p.d_pred = d_pred;
  p.n = n; p.C = C; p.width = width; p.height = height;
  ynb::rgb_loss_bwd_kernel<<<(unsigned)((B * n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, B);
  return ynb::check_launch("yn_rgb_loss_bwd");
}

// ------------------------------------------------------------------------------------------------
// Rasterised Monte-Carlo samples (`scatter_rays_to_image`, pipelines/utils.py:299-323, called for rgb / depth / alpha
// every training step when `output_rasterized_mc` is set, nerf_pipeline.py:307-324): the per-ray values of up to three
// tensors are written to their (pre-zeroed) [B, H, W, C_k] canvases in ONE launch.
// ------------------------------------------------------------------------------------------------
namespace ynb {
struct ScatterParams {
  const float* src[3];
  float* dst[3];
  int C[3];
  int n_tensors;
  const float* xy;  // [B, n, 2]
  int64_t B, n;
  int width, height;
};
__global__ void __launch_bounds__(256) scatter_rays_kernel(const ScatterParams p) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p.B * p.n) return;
  const int64_t b = t / p.n;
  const float x = __ldg(p.xy + t * 2), y = __ldg(p.xy + t * 2 + 1);
  const int64_t pix = (int64_t)__fadd_rn(x, __fmul_rn((float)p.width, y));
  const int64_t n_pix = (int64_t)p.width * p.height;
  for (int k = 0; k < p.n_tensors; ++k)
    for (int c = 0; c < p.C[k]; ++c) p.dst[k][(b * n_pix + pix) * p.C[k] + c] = __ldg(p.src[k] + t * p.C[k] + c);
}
}  // namespace ynb

extern "C" int yn_scatter_rays(const float* const* src, float* const* dst, const int* channels, int n_tensors, const float* xy,
                               int64_t B, int64_t n, int width, int height, void* stream) {
  if (B < 0 || n < 0 || n_tensors < 1 || n_tensors > 3 || width < 1 || height < 1)
    return ynb::fail(YN_ERR_INVALID_ARGUMENT, "yn_scatter_rays: bad sizes");
  if (B * n == 0) return YN_OK;
  if (!src || !dst || !channels || !xy) return ynb::fail(YN_ERR_INVALID_ARGUMENT, "yn_scatter_rays: null pointer");
  ynb::ScatterParams p = {};
  for (int k = 0; k < n_tensors; ++k) {
    if (!src[k] || !dst[k] || channels[k] < 1) return ynb::fail(YN_ERR_INVALID_ARGUMENT, "yn_scatter_rays: null tensor");
    p.src[k] = src[k]; p.dst[k] = dst[k]; p.C[k] = channels[k];
  }
  p.n_tensors = n_tensors; p.xy = xy; p.B = B; p.n = n; p.width = width; p.height = height;
  ynb::scatter_rays_kernel<<<(unsigned)((B * n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return ynb::check_launch("yn_scatter_rays");
}
