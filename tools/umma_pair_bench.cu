// Microbenchmark (not part of the product library): tcgen05.mma throughput of the MLP's GEMM shape with
//   mode 1: cta_group::1, M=128 N=128 (what mlp_fwd_kernel issues today), one CTA per SM
//   mode 2: cta_group::2, M=256 N=256 on an SM pair (each SM supplies its 128 rows of A and 128 of the 256 rows of B)
// A tile (128 points x 256 features, 4 swizzled blocks) stays resident in shared memory; weight blocks
// [128 n x 64 k] stream through a 4-deep ring with TMA bulk copies, exactly like the forward kernel; there is no
// epilogue, so the number is the MMA + operand-feed ceiling.  The last layer's accumulator is written out and checked.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o umma_pair_bench umma_pair_bench.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../yet-another-nerf_b200/csrc/sm100_ptx.cuh"

using namespace ynb;

constexpr int kBlk = 16384;
constexpr int kRing = 4;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra D_%=;\n\t"
      "bra W_%=;\n\t"
      "D_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}

struct Params {
  const uint8_t* a_tiles;  // [n_tiles][4 blocks] swizzled fp16
  const uint8_t* w_blocks; // [2 halves][4 kb][16 KB]: half h = output features h*128..h*128+127
  float* out;              // [n_tiles][128][256]
  int n_tiles;
  int layers;
};

// ------------------------------------------------------------------ mode 1: what the product kernel does today
__global__ void __launch_bounds__(192, 1) bench_1cta(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_a = base, s_ring = base + 4 * kBlk, s_bar = s_ring + kRing * kBlk;
  const uint32_t bar_full = s_bar, bar_empty = s_bar + 32, bar_a = s_bar + 64, bar_done = s_bar + 72, s_tptr = s_bar + 80;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kRing; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    mbar_init(bar_a, 1); mbar_init(bar_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) { tmem_alloc(s_tptr, 256); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t tmem; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(s_tptr));
  const int tile = blockIdx.x;
  if (warp == 0 && elect_one()) {
    mbar_arrive_expect_tx(bar_a, 4 * kBlk);
    bulk_g2s(s_a, p.a_tiles + (size_t)tile * 4 * kBlk, 4 * kBlk, bar_a);
    uint32_t slot = 0, ph = 0;
    for (int l = 0; l < p.layers; ++l)
      for (int s = 0; s < 8; ++s) {  // nh-major, kb-minor like the product kernel
        mbar_wait(bar_empty + 8 * slot, ph ^ 1);
        mbar_arrive_expect_tx(bar_full + 8 * slot, kBlk);
        bulk_g2s(s_ring + slot * kBlk, p.w_blocks + (size_t)s * kBlk, kBlk, bar_full + 8 * slot);
        if (++slot == kRing) { slot = 0; ph ^= 1; }
      }
  } else if (warp == 1 && elect_one()) {
    constexpr uint32_t idesc = umma_idesc(128, 128, 0, 0, 0);
    mbar_wait(bar_a, 0);
    uint32_t slot = 0, ph = 0;
    for (int l = 0; l < p.layers; ++l)
      for (int nh = 0; nh < 2; ++nh)
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(bar_full + 8 * slot, ph);
          tc_fence_after();
          const uint64_t a = umma_desc_kmajor(s_a + kb * kBlk), b = umma_desc_kmajor(s_ring + slot * kBlk);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem + nh * 128, a + 2 * k, b + 2 * k, idesc, (kb | k) != 0);
          umma_commit(bar_empty + 8 * slot);
          if (++slot == kRing) { slot = 0; ph ^= 1; }
        }
    umma_commit(bar_done);
  }
  __syncwarp();
  if (warp >= 2) {
    mbar_wait(bar_done, 0);
    tc_fence_after();
    const int q = warp & 3, row = q * 32 + lane;
    float* o = p.out + ((size_t)tile * 128 + row) * 256;
    for (int cb = 0; cb < 8; ++cb) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + cb * 32, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) o[cb * 32 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------ mode 2: CTA pair
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1) bench_2cta(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_a = base, s_ring = base + 4 * kBlk, s_bar = s_ring + kRing * kBlk;
  // full[4] (local TMA), empty[4] (MMA done, multicast), peer_full[4] (leader only: the peer's block landed),
  // a_full, peer_a, done, tmem ptr
  const uint32_t bar_full = s_bar, bar_empty = s_bar + 32, bar_pfull = s_bar + 64, bar_a = s_bar + 96, bar_pa = s_bar + 104,
                 bar_done = s_bar + 112, s_tptr = s_bar + 120;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    for (int i = 0; i < kRing; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); mbar_init(bar_pfull + 8 * i, 1); }
    mbar_init(bar_a, 1); mbar_init(bar_pa, 1); mbar_init(bar_done, 1);
    mbar_fence_init();
  }
  cluster_sync_all();
  if (warp == 1) { tmem_alloc2(s_tptr, 256); tmem_relinquish2(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  cluster_sync_all();
  uint32_t tmem; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(s_tptr));
  const int tile = blockIdx.x;  // each CTA of the pair owns one 128-point tile
  if (warp == 0 && elect_one()) {
    mbar_arrive_expect_tx(bar_a, 4 * kBlk);
    bulk_g2s(s_a, p.a_tiles + (size_t)tile * 4 * kBlk, 4 * kBlk, bar_a);
    uint32_t slot = 0, ph = 0;
    for (int l = 0; l < p.layers; ++l)
      for (int kb = 0; kb < 4; ++kb) {  // this CTA's half of the 256 output features
        mbar_wait(bar_empty + 8 * slot, ph ^ 1);
        mbar_arrive_expect_tx(bar_full + 8 * slot, kBlk);
        bulk_g2s(s_ring + slot * kBlk, p.w_blocks + ((size_t)rank * 4 + kb) * kBlk, kBlk, bar_full + 8 * slot);
        if (++slot == kRing) { slot = 0; ph ^= 1; }
      }
  } else if (warp == 1 && elect_one()) {
    if (rank == 1) {
      // forwarder: tell the leader that this CTA's operands have landed
      mbar_wait(bar_a, 0);
      mbar_arrive_cluster(bar_pa, 0);
      uint32_t slot = 0, ph = 0;
      for (int l = 0; l < p.layers; ++l)
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(bar_full + 8 * slot, ph);
          mbar_arrive_cluster(bar_pfull + 8 * slot, 0);
          if (++slot == kRing) { slot = 0; ph ^= 1; }
        }
    } else {
      constexpr uint32_t idesc = umma_idesc(256, 256, 0, 0, 0);
      mbar_wait(bar_a, 0);
      mbar_wait_cluster(bar_pa, 0);
      uint32_t slot = 0, ph = 0;
      for (int l = 0; l < p.layers; ++l)
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(bar_full + 8 * slot, ph);
          mbar_wait_cluster(bar_pfull + 8 * slot, ph);
          tc_fence_after();
          const uint64_t a = umma_desc_kmajor(s_a + kb * kBlk), b = umma_desc_kmajor(s_ring + slot * kBlk);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_2cta(tmem, a + 2 * k, b + 2 * k, idesc, (kb | k) != 0);
          umma_commit_2cta(bar_empty + 8 * slot, 3);
          if (++slot == kRing) { slot = 0; ph ^= 1; }
        }
      umma_commit_2cta(bar_done, 3);
    }
  }
  __syncwarp();
  if (warp >= 2) {
    mbar_wait(bar_done, 0);
    tc_fence_after();
    const int q = warp & 3, row = q * 32 + lane;
    float* o = p.out + ((size_t)tile * 128 + row) * 256;
    for (int cb = 0; cb < 8; ++cb) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + cb * 32, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) o[cb * 32 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before(); __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc2(tmem, 256);
}

static void pack_block(std::vector<uint8_t>& dst, size_t off, const std::vector<float>& src, int ld, int row0, int col0) {
  for (int r = 0; r < 128; ++r)
    for (int c = 0; c < 64; ++c) {
      __half h = __float2half(src[(size_t)(row0 + r) * ld + col0 + c]);
      *reinterpret_cast<__half*>(&dst[off + sw128_offset(r, c)]) = h;
    }
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 1;
  const int layers = argc > 2 ? atoi(argv[2]) : 2000;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int n_tiles = sms;  // one tile per SM
  std::vector<float> A((size_t)n_tiles * 128 * 256), W(256 * 256);
  srand(1);
  for (auto& x : A) x = (rand() % 2001 - 1000) / 1000.f;
  for (auto& x : W) x = (rand() % 2001 - 1000) / 4000.f;
  std::vector<uint8_t> a_img((size_t)n_tiles * 4 * kBlk), w_img(8 * kBlk);
  for (int t = 0; t < n_tiles; ++t)
    for (int kb = 0; kb < 4; ++kb) pack_block(a_img, ((size_t)t * 4 + kb) * kBlk, A, 256, t * 128, kb * 64);
  for (int nh = 0; nh < 2; ++nh)
    for (int kb = 0; kb < 4; ++kb) pack_block(w_img, ((size_t)nh * 4 + kb) * kBlk, W, 256, nh * 128, kb * 64);
  uint8_t *d_a, *d_w;
  float* d_o;
  cudaMalloc(&d_a, a_img.size());
  cudaMalloc(&d_w, w_img.size());
  cudaMalloc(&d_o, (size_t)n_tiles * 128 * 256 * 4);
  cudaMemcpy(d_a, a_img.data(), a_img.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(d_w, w_img.data(), w_img.size(), cudaMemcpyHostToDevice);
  Params p{d_a, d_w, d_o, n_tiles, layers};
  const int smem = 8 * kBlk + 256 + 1024;
  cudaFuncSetAttribute(bench_1cta, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(bench_2cta, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int it = 0; it < 4; ++it) {
    cudaEventRecord(e0);
    if (mode == 1) bench_1cta<<<n_tiles, 192, smem>>>(p);
    else bench_2cta<<<n_tiles, 192, smem>>>(p);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); return 1; }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double flops = 2.0 * n_tiles * 128.0 * 256 * 256 * layers;
  std::vector<float> out((size_t)n_tiles * 128 * 256);
  cudaMemcpy(out.data(), d_o, out.size() * 4, cudaMemcpyDeviceToHost);
  double max_err = 0;
  for (int t : {0, 1, n_tiles - 1})
    for (int r = 0; r < 128; r += 17)
      for (int n = 0; n < 256; n += 5) {
        double ref = 0;
        for (int k = 0; k < 256; ++k)
          ref += (double)__half2float(__float2half(A[((size_t)t * 128 + r) * 256 + k])) * __half2float(__float2half(W[(size_t)n * 256 + k]));
        max_err = fmax(max_err, fabs(ref - out[((size_t)t * 128 + r) * 256 + n]));
      }
  printf("{\"mode\": %d, \"layers\": %d, \"ms\": %.3f, \"tflops\": %.1f, \"max_err\": %.3e}\n", mode, layers, best, flops / best / 1e9, max_err);
  return 0;
}
