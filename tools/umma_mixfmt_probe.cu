// Probe (not part of the product library): does tcgen05.mma.kind::f16 accept A = bf16 with B = fp16 (and the
// reverse) in one instruction?  The instruction descriptor has separate A / B format fields.  One CTA, one
// M128 x N128 x K64 product, checked against a double-precision reference of the rounded operands.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o umma_mixfmt_probe umma_mixfmt_probe.cu
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../yet-another-nerf_b200/csrc/sm100_ptx.cuh"

using namespace ynb;
constexpr int kBlk = 16384;

__global__ void __launch_bounds__(192, 1) probe(const uint8_t* a, const uint8_t* b, float* out, uint32_t idesc) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_a = base, s_b = base + kBlk, s_bar = s_b + kBlk, bar_done = s_bar + 8, s_tptr = s_bar + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(s_bar, 1); mbar_init(bar_done, 1); mbar_fence_init(); }
  if (warp == 1) { tmem_alloc(s_tptr, 128); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t tmem; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(s_tptr));
  if (warp == 0 && elect_one()) {
    mbar_arrive_expect_tx(s_bar, 2 * kBlk);
    bulk_g2s(s_a, a, kBlk, s_bar);
    bulk_g2s(s_b, b, kBlk, s_bar);
  } else if (warp == 1 && elect_one()) {
    mbar_wait(s_bar, 0);
    tc_fence_after();
    const uint64_t ad = umma_desc_kmajor(s_a), bd = umma_desc_kmajor(s_b);
    for (int k = 0; k < 4; ++k) umma_f16(tmem, ad + 2 * k, bd + 2 * k, idesc, k != 0);
    umma_commit(bar_done);
  }
  __syncwarp();
  if (warp >= 2) {
    mbar_wait(bar_done, 0);
    tc_fence_after();
    const int q = warp & 3, row = q * 32 + lane;
    for (int cb = 0; cb < 4; ++cb) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + cb * 32, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) out[row * 128 + cb * 32 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 128);
}

static float rnd(float x, int fmt) { return fmt == 1 ? __bfloat162float(__float2bfloat16(x)) : __half2float(__float2half(x)); }
static void pack(std::vector<uint8_t>& dst, const std::vector<float>& src, int fmt) {
  for (int r = 0; r < 128; ++r)
    for (int c = 0; c < 64; ++c) {
      const float x = src[r * 64 + c];
      if (fmt == 1) *reinterpret_cast<__nv_bfloat16*>(&dst[sw128_offset(r, c)]) = __float2bfloat16(x);
      else *reinterpret_cast<__half*>(&dst[sw128_offset(r, c)]) = __float2half(x);
    }
}

int main() {
  std::vector<float> A(128 * 64), B(128 * 64);
  srand(3);
  // values whose fp16 and bf16 roundings differ visibly; a few tiny ones that fp16 flushes towards zero
  for (auto& x : A) x = (rand() % 20001 - 10000) / 9973.f;
  for (auto& x : B) x = (rand() % 20001 - 10000) / 7919.f;
  A[5] = 3.1e-7f; A[70] = -2.7e-8f;
  uint8_t *da, *db; float* dout;
  cudaMalloc(&da, kBlk); cudaMalloc(&db, kBlk); cudaMalloc(&dout, 128 * 128 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kBlk + 2048);
  for (int af = 0; af < 2; ++af)
    for (int bf = 0; bf < 2; ++bf) {
      std::vector<uint8_t> ia(kBlk), ib(kBlk);
      pack(ia, A, af); pack(ib, B, bf);
      cudaMemcpy(da, ia.data(), kBlk, cudaMemcpyHostToDevice);
      cudaMemcpy(db, ib.data(), kBlk, cudaMemcpyHostToDevice);
      cudaMemset(dout, 0, 128 * 128 * 4);
      const uint32_t idesc = (1u << 4) | ((uint32_t)af << 7) | ((uint32_t)bf << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
      probe<<<1, 192, 2 * kBlk + 2048>>>(da, db, dout, idesc);
      cudaError_t err = cudaDeviceSynchronize();
      if (err != cudaSuccess) { printf("{\"a_fmt\": %d, \"b_fmt\": %d, \"error\": \"%s\"}\n", af, bf, cudaGetErrorString(err)); return 1; }
      std::vector<float> out(128 * 128);
      cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
      double max_err = 0, max_ref = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 128; ++n) {
          double ref = 0;
          for (int k = 0; k < 64; ++k) ref += (double)rnd(A[m * 64 + k], af) * rnd(B[n * 64 + k], bf);
          max_err = fmax(max_err, fabs(ref - out[m * 128 + n]));
          max_ref = fmax(max_ref, fabs(ref));
        }
      printf("{\"a_fmt\": %d, \"b_fmt\": %d, \"max_err\": %.3e, \"max_ref\": %.3f}\n", af, bf, max_err, max_ref);
    }
  return 0;
}
