// Ray sampler kernels: pixel coordinates -> origins / directions / depths with optional stratified jitter.
// Replaces _xy_to_ray_bundle (yanerf/pipelines/ray_samplers/ray_sampler.py:249-314) and
// _jiggle_within_stratas (ray_sampler.py:361-386).  Pure bandwidth; one thread per (ray, sample).
#include <cuda_runtime.h>

#include "mlp_common.cuh"
#include "rng.cuh"

namespace ynb {

struct RayParams {
  const float* poses;
  int64_t pose_bs, pose_rs;
  const float* focal;
  const float* xy;
  const float* depths;
  const float* u;
  RngRef rng;            // stratified jitter drawn in the kernel when `u` is null and `stratified` is set
  int stratified;
  const int64_t* pick;   // != nullptr: the pixel of ray i is picked here (keyed Feistel permutation), `xy` is ignored
  int64_t* idx_out;      // picked flat pixel indices [B, n] (optional)
  int half_bits;
  float* origins;
  float* directions;
  float* lengths;
  float* xy_out;
  int64_t B, n;
  int P, width, height, full_grid;
};


__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

// i-th element of a keyed pseudo-random permutation of [0, n_pix): 6-round Feistel network on 2 * half_bits bits with
// cycle walking (expected < 4 rounds: 2^(2 * half_bits) < 4 * n_pix).  Distinct i give distinct pixels.
__device__ __forceinline__ uint64_t feistel_pick(uint64_t seed0, int64_t b, uint64_t i, uint64_t n_pix, int half_bits) {
  const uint64_t seed = seed0 + 0x9E3779B97F4A7C15ULL * (uint64_t)(b + 1);
  const uint32_t mask = (1u << half_bits) - 1u;
  uint64_t v = i;
  do {
    uint32_t l = (uint32_t)(v >> half_bits) & mask, r = (uint32_t)v & mask;
#pragma unroll
    for (int round = 0; round < 6; ++round) {
      const uint32_t key = (uint32_t)(seed >> (8 * (round & 3))) + 0x85ebca6bU * (round + 1);
      const uint32_t f = mix32(r ^ key) & mask;
      const uint32_t nl = r;
      r = l ^ f;
      l = nl;
    }
    v = ((uint64_t)l << half_bits) | r;
  } while (v >= n_pix);
  return v;
}

__global__ void __launch_bounds__(256) ray_bundle_kernel(const RayParams p) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over B*n*P
  const int64_t total = p.B * p.n * p.P;
  if (idx >= total) return;
  const int s = (int)(idx % p.P);
  const int64_t ray = idx / p.P;
  const int64_t b = ray / p.n, i = ray % p.n;
  // depths: linspace (host-computed, torch.linspace on CPU) + stratified jitter
  const float z = __ldg(p.depths + s);
  float out = z;
  if (p.u != nullptr || p.stratified) {
    // mids = 0.5 * (z[1:] + z[:-1]); lower = cat(z[:1], mids); upper = cat(mids, z[-1:])
    const float lower = s == 0 ? z : __fmul_rn(0.5f, __fadd_rn(z, __ldg(p.depths + s - 1)));
    const float upper = s == p.P - 1 ? z : __fmul_rn(0.5f, __fadd_rn(__ldg(p.depths + s + 1), z));
    const float u = p.u != nullptr ? __ldg(p.u + idx) : UniformRow(p.rng, ray).get(s);
    out = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), u));
  }
  p.lengths[idx] = out;
  if (s == 0) {
    float x, y;
    if (p.pick != nullptr) {
      const uint64_t seed = (uint64_t)p.pick[0] ^ (0xD1B54A32D192ED03ULL * (uint64_t)(p.pick[2] + 1));
      const uint64_t v = feistel_pick(seed, b, (uint64_t)i, (uint64_t)p.width * p.height, p.half_bits);
      if (p.idx_out) p.idx_out[ray] = (int64_t)v;
      x = (float)(v % p.width);
      y = (float)(v / p.width);
    } else if (p.full_grid) {
      x = (float)(i % p.width);
      y = (float)(i / p.width);
    } else {
      x = __ldg(p.xy + ray * 2);
      y = __ldg(p.xy + ray * 2 + 1);
    }
    if (p.xy_out) {
      p.xy_out[ray * 2] = x;
      p.xy_out[ray * 2 + 1] = y;
    }
    const float* pose = p.poses + b * p.pose_bs;
    const float f = __ldg(p.focal + b);
    const float cx = __fdiv_rn(__fsub_rn(x, (float)(p.width * 0.5)), f);
    const float cy = __fdiv_rn(__fsub_rn(y, (float)(p.height * 0.5)), f);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float* row = pose + r * p.pose_rs;
      const float d = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(row), cx), __fmul_rn(__ldg(row + 1), cy)), __ldg(row + 2));
      p.directions[ray * 3 + r] = d;
      p.origins[ray * 3 + r] = __ldg(row + 3);
    }
  }
}

}  // namespace ynb

extern "C" int yn_ray_bundle(const float* poses, int64_t pose_batch_stride, int64_t pose_row_stride,
                             const float* focal, const float* xy, const float* depths, const float* u,
                             float* origins, float* directions, float* lengths, float* xy_out, int64_t B, int64_t n,
                             int P, int width, int height, int full_grid, void* stream) {
  if (B < 0 || n < 0 || P < 1 || width < 1 || height < 1)
    return ynb::fail(YN_ERR_INVALID_ARGUMENT, "yn_ray_bundle: bad sizes");
  if (B * n == 0) return YN_OK;
  if (!poses || !focal || !depths || !origins || !directions || !lengths || (!xy && !full_grid))
    return ynb::fail(YN_ERR_INVALID_ARGUMENT, "yn_ray_bundle: null pointer");
  ynb::RayParams p;
  p.poses = poses; p.pose_bs = pose_batch_stride; p.pose_rs = pose_row_stride; p.focal = focal; p.xy = xy;
  p.depths = depths; p.u = u; p.rng.state = nullptr; p.rng.site = 0; p.stratified = 0; p.pick = nullptr; p.idx_out = nullptr;
  p.half_bits = 0; p.origins = origins; p.directions = directions; p.lengths = lengths; p.xy_out = xy_out;
  p.B = B; p.n = n; p.P = P; p.width = width; p.height = height; p.full_grid = full_grid;
  const int64_t total = B * n * P;
  ynb::ray_bundle_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return ynb::check_launch("yn_ray_bundle");
}

// ------------------------------------------------------------------------------------------------
// Pixel pick for training: n DISTINCT pixels per image, uniformly at random (what
// `torch.multinomial(ones(H*W), n, replacement=False)` draws in _RaySampler.forward, ray_sampler.py:187-229, for the
// unmasked case).  A keyed Feistel permutation of [0, 2^k) with cycle walking maps i = 0..n-1 to distinct pixels in
// O(n) -- no pass over the H*W weights.  The 64-bit seed is read from device memory so that a captured CUDA graph
// draws new pixels on every replay (the runner bumps it on the stream).
// ------------------------------------------------------------------------------------------------
namespace ynb {

__global__ void __launch_bounds__(256) sample_pixels_kernel(const int64_t* __restrict__ seed_ptr, int64_t* __restrict__ idx,
                                                           float* __restrict__ xy, int64_t B, int64_t n, int64_t n_pix,
                                                           int width, int half_bits) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * n) return;
  const uint64_t v = feistel_pick((uint64_t)seed_ptr[0], t / n, (uint64_t)(t % n), (uint64_t)n_pix, half_bits);
  idx[t] = (int64_t)v;
  if (xy) {
    xy[2 * t] = (float)(v % width);
    xy[2 * t + 1] = (float)(v / width);
  }
}

// One thread at the start of every training step: the step's draws are keyed by a SNAPSHOT of the step counter, so every
// kernel of the step (a forward kernel and its backward in particular) regenerates the same values, and a replayed CUDA
// graph advances on its own.  The Adam step counter (float, yn_adam_step_dev) is bumped by the same launch.
__global__ void step_begin_kernel(int64_t* rng_state, float* adam_state) {
  if (rng_state != nullptr) {
    rng_state[2] = rng_state[1];
    rng_state[1] = rng_state[1] + 1;
  }
  if (adam_state != nullptr) adam_state[0] += 1.f;
}

}  // namespace ynb

extern "C" int yn_sample_pixels(const int64_t* seed, int64_t* idx, float* xy, int64_t B, int64_t n, int width, int height,
                                void* stream) {
  const int64_t n_pix = (int64_t)width * height;
  if (B < 0 || n < 0 || width < 1 || height < 1 || n > n_pix)
    return ynb::fail(YN_ERR_INVALID_ARGUMENT, "yn_sample_pixels: need 0 <= n <= width*height");
  if (B * n == 0) return YN_OK;
  if (!seed || !idx) return ynb::fail(YN_ERR_INVALID_ARGUMENT, "yn_sample_pixels: null pointer");
  int half_bits = 1;
  while (((int64_t)1 << (2 * half_bits)) < n_pix) ++half_bits;
  ynb::sample_pixels_kernel<<<(unsigned)((B * n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      seed, idx, xy, B, n, n_pix, width, half_bits);
  return ynb::check_launch("yn_sample_pixels");
}

extern "C" int yn_step_begin(int64_t* rng_state, float* adam_state, void* stream) {
  if (!rng_state && !adam_state) return YN_OK;
  ynb::step_begin_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(rng_state, adam_state);
  return ynb::check_launch("yn_step_begin");
}

// Training rays in ONE launch: pixel pick (the unmasked multinomial of ray_sampler.py:187-229), pinhole rays
// (249-314) and stratified depths (361-386) with the jitter drawn in the kernel from `rng_state` (site `rng_site`).
extern "C" int yn_train_rays(const int64_t* rng_state, int rng_site, const float* poses, int64_t pose_batch_stride,
                             int64_t pose_row_stride, const float* focal, const float* depths, int stratified,
                             int64_t* idx_out, float* xy_out, float* origins, float* directions, float* lengths, int64_t B,
                             int64_t n, int P, int width, int height, void* stream) {
  const int64_t n_pix = (int64_t)width * height;
  if (B < 0 || n < 0 || P < 1 || width < 1 || height < 1 || n > n_pix)
    return ynb::fail(YN_ERR_INVALID_ARGUMENT, "yn_train_rays: need 0 <= n <= width*height, P >= 1");
  if (B * n == 0) return YN_OK;
  if (!rng_state || !poses || !focal || !depths || !xy_out || !origins || !directions || !lengths)
    return ynb::fail(YN_ERR_INVALID_ARGUMENT, "yn_train_rays: null pointer");
  ynb::RayParams p;
  p.poses = poses; p.pose_bs = pose_batch_stride; p.pose_rs = pose_row_stride; p.focal = focal; p.xy = nullptr;
  p.depths = depths; p.u = nullptr; p.rng.state = rng_state; p.rng.site = rng_site; p.stratified = stratified;
  p.pick = rng_state; p.idx_out = idx_out; p.origins = origins; p.directions = directions; p.lengths = lengths;
  p.xy_out = xy_out; p.B = B; p.n = n; p.P = P; p.width = width; p.height = height; p.full_grid = 0;
  p.half_bits = 1;
  while (((int64_t)1 << (2 * p.half_bits)) < n_pix) ++p.half_bits;
  const int64_t total = B * n * P;
  ynb::ray_bundle_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return ynb::check_launch("yn_train_rays");
}

// The draws the kernels generate for (site, current step), written out: out[row, e] for row < R, e < P.  kind 0 = U[0,1),
// 1 = N(0,1).  Test / debugging aid: feeding these back through the explicit draw pointers must reproduce the in-kernel
// path bit for bit.
namespace ynb {
__global__ void __launch_bounds__(256) rng_fill_kernel(RngRef r, float* out, int64_t R, int P, int kind) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= R * P) return;
  const int64_t row = t / P;
  const int e = (int)(t % P);
  out[t] = kind == 0 ? UniformRow(r, row).get(e) : NormalRow(r, row).get(e);
}
}  // namespace ynb

extern "C" int yn_rng_fill(const int64_t* rng_state, int rng_site, int kind, float* out, int64_t R, int P, void* stream) {
  if (R < 0 || P < 1 || (kind != 0 && kind != 1)) return ynb::fail(YN_ERR_INVALID_ARGUMENT, "yn_rng_fill: bad arguments");
  if (R == 0) return YN_OK;
  if (!rng_state || !out) return ynb::fail(YN_ERR_INVALID_ARGUMENT, "yn_rng_fill: null pointer");
  ynb::RngRef r;
  r.state = rng_state;
  r.site = rng_site;
  ynb::rng_fill_kernel<<<(unsigned)((R * P + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(r, out, R, P, kind);
  return ynb::check_launch("yn_rng_fill");
}
