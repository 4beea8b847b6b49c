// Counter-based random draws for the training step: Philox4x32-10 keyed by a 64-bit seed, counter =
// (row, group of four consecutive elements, draw site, step).  Every draw of the reference's training forward --
// stratified jitter `rand_like` (ray_sampler.py:384), density noise `randn_like`
// (multipass_emission_absorpsion_renderer.py:203-207), inverse-CDF uniforms `torch.rand` (renderers/utils.py:133-134)
// -- can be produced inside the consuming kernel instead of round-tripping an [R, P] tensor through HBM; an explicit
// draw pointer (the parity tests' replayed draws) always takes precedence.
//
// Device state `int64 state[4]`: [0] seed, [1] number of steps begun, [2] the CURRENT step (snapshot written by
// yn_step_begin, constant while a step's kernels run, so a forward kernel and its backward regenerate the same noise),
// [3] reserved.
#pragma once
#include <stdint.h>

namespace ynb {

struct RngRef {
  const int64_t* state;  // nullptr: no in-kernel draws
  int site;              // distinguishes the draw sites of one step (0..255)
};

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

// the four 32-bit words of (row, group)
__device__ __forceinline__ uint4 rng_words(const RngRef& r, int64_t row, uint32_t group) {
  const uint64_t seed = static_cast<uint64_t>(r.state[0]), step = static_cast<uint64_t>(r.state[2]);
  const uint4 ctr = make_uint4(static_cast<uint32_t>(row), (static_cast<uint32_t>(static_cast<uint64_t>(row) >> 32) << 8) | (r.site & 255),
                               group, static_cast<uint32_t>(step));
  const uint2 key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32) + static_cast<uint32_t>(step >> 32));
  return philox4x32_10(ctr, key);
}

// U[0,1) with 24 random bits, like torch.rand for float32
__device__ __forceinline__ float rng_u01(uint32_t x) { return static_cast<float>(x >> 8) * 0x1p-24f; }

// two N(0,1) values from two words (Box-Muller; the radius uses (0,1] so the logarithm is finite)
__device__ __forceinline__ float2 rng_normal2(uint32_t a, uint32_t b) {
  const float u1 = static_cast<float>((a >> 8) + 1u) * 0x1p-24f;
  const float u2 = static_cast<float>(b >> 8) * 0x1p-24f;
  const float rad = sqrtf(-2.f * __logf(u1));
  float s, c;
  __sincosf(6.283185307179586f * u2, &s, &c);
  return make_float2(rad * c, rad * s);
}

__device__ __forceinline__ float sel4(const float (&v)[4], int i) {
  return i == 0 ? v[0] : (i == 1 ? v[1] : (i == 2 ? v[2] : v[3]));
}

// Element e of a row -> group e / 4, component e % 4.  Caches the last group so that a lane walking consecutive
// elements evaluates Philox once per four of them.
struct UniformRow {
  RngRef r;
  int64_t row;
  int group;
  float v[4];
  __device__ __forceinline__ UniformRow(const RngRef& ref, int64_t row_) : r(ref), row(row_), group(-1) {}
  __device__ __forceinline__ float get(int e) {
    const int g = e >> 2;
    if (g != group) {
      const uint4 w = rng_words(r, row, static_cast<uint32_t>(g));
      v[0] = rng_u01(w.x); v[1] = rng_u01(w.y); v[2] = rng_u01(w.z); v[3] = rng_u01(w.w);
      group = g;
    }
    return sel4(v, e & 3);
  }
};
struct NormalRow {
  RngRef r;
  int64_t row;
  int group;
  float v[4];
  __device__ __forceinline__ NormalRow(const RngRef& ref, int64_t row_) : r(ref), row(row_), group(-1) {}
  __device__ __forceinline__ float get(int e) {
    const int g = e >> 2;
    if (g != group) {
      const uint4 w = rng_words(r, row, static_cast<uint32_t>(g));
      const float2 a = rng_normal2(w.x, w.y), b = rng_normal2(w.z, w.w);
      v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
      group = g;
    }
    return sel4(v, e & 3);
  }
};

}  // namespace ynb
