// Microbenchmark (not part of the product library): what the L2 <-> SM path and HBM sustain for the 16 KB
// bulk copies (cp.async.bulk) the MLP kernels are built from, one persistent CTA per SM.
// Streams (each optional): R1 = read stream, W = write stream, R2 = second read stream.
// Regions: "own" = every CTA cycles over its own 256 KB (37 MB in total, L2 resident), "nbr" = the 256 KB region of
// the next CTA (what a consumer SM reads from a producer SM's ring), "shared" = all CTAs cycle over the same 2.4 MB
// (weight stream), "hbm" = every CTA streams through its own slice of 8 GB.
// hint: 0 none, 1 = L2::evict_last on the ring streams and L2::evict_first on the hbm streams.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o l2_bw_bench l2_bw_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "../yet-another-nerf_b200/csrc/sm100_ptx.cuh"

using namespace ynb;

constexpr int kBlk = 16384;
constexpr int kSlots = 5;

struct Stream {
  uint8_t* base;
  size_t span, stride;  // bytes each CTA cycles over; distance between the CTAs' regions
  int shift;            // region of CTA (blockIdx + shift) % grid
  int n;                // blocks per CTA
  int hint;             // 0 none, 1 evict_last, 2 evict_first
};
struct Params {
  Stream r1, w, r2;
};

__device__ __forceinline__ uint64_t make_policy(int hint) {
  uint64_t pol = 0;
  if (hint == 1) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  if (hint == 2) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_s2g_hint(void* dst, uint32_t src, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(src), "r"(bytes), "l"(pol)
               : "memory");
}

__device__ void read_stream(const Stream& s, uint32_t s_buf, uint32_t s_bar) {
  const uint8_t* src = s.base + (size_t)((blockIdx.x + s.shift) % gridDim.x) * s.stride;
  const size_t nblk = s.span / kBlk;
  const uint64_t pol = make_policy(s.hint);
  for (int i = 0; i < s.n + kSlots; ++i) {
    const int slot = i % kSlots;
    if (i >= kSlots) mbar_wait(s_bar + 8 * slot, ((i / kSlots) - 1) & 1);
    if (i < s.n) {
      mbar_arrive_expect_tx(s_bar + 8 * slot, kBlk);
      const uint8_t* a = src + ((size_t)i % nblk) * kBlk;
      if (s.hint) bulk_g2s_hint(s_buf + slot * kBlk, a, kBlk, s_bar + 8 * slot, pol);
      else bulk_g2s(s_buf + slot * kBlk, a, kBlk, s_bar + 8 * slot);
    }
  }
}

__global__ void __launch_bounds__(96, 1) bw_kernel(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_r1 = base, s_r2 = base + kSlots * kBlk, s_wr = s_r2 + kSlots * kBlk, s_bar = s_wr + 2 * kBlk;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * kSlots; ++i) mbar_init(s_bar + 8 * i, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0 && p.r1.n > 0) read_stream(p.r1, s_r1, s_bar);
  if (threadIdx.x == 64 && p.r2.n > 0) read_stream(p.r2, s_r2, s_bar + 8 * kSlots);
  if (threadIdx.x == 32 && p.w.n > 0) {
    uint8_t* dst = p.w.base + (size_t)((blockIdx.x + p.w.shift) % gridDim.x) * p.w.stride;
    const size_t nblk = p.w.span / kBlk;
    const uint64_t pol = make_policy(p.w.hint);
    for (int i = 0; i < p.w.n; ++i) {
      uint8_t* a = dst + ((size_t)i % nblk) * kBlk;
      if (p.w.hint) bulk_s2g_hint(a, s_wr + (i & 1) * kBlk, kBlk, pol);
      else bulk_s2g(a, s_wr + (i & 1) * kBlk, kBlk);
      bulk_commit();
      bulk_wait_read<6>();
    }
    bulk_wait<0>();
  }
}

int main(int argc, char** argv) {
  const int n_blocks = argc > 1 ? atoi(argv[1]) : 40000;  // 16 KB blocks per CTA and stream
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int n_ctas = argc > 2 ? atoi(argv[2]) : sms;  // fewer CTAs than SMs: is the cap per SM or chip-wide?
  const size_t big = (size_t)8 << 30;
  uint8_t *d_a, *d_b, *d_ring;
  cudaMalloc(&d_a, big);
  cudaMalloc(&d_b, big);
  cudaMalloc(&d_ring, (size_t)64 << 20);
  cudaMemset(d_a, 1, big);
  cudaMemset(d_b, 0, big);
  cudaMemset(d_ring, 0, (size_t)64 << 20);
  const int smem = (2 * kSlots + 2) * kBlk + 256 + 1024;
  cudaFuncSetAttribute(bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const size_t small = 256 << 10, shared_w = (2400 << 10) / kBlk * kBlk, slice = (big / sms) / kBlk * kBlk;
  const Stream none{nullptr, small, small, 0, 0, 0};
  auto own = [&](int hint) { return Stream{d_ring, small, small, 0, n_blocks, hint}; };
  auto nbr = [&](int hint) { return Stream{d_ring, small, small, 1, n_blocks, hint}; };
  auto shared = [&]() { return Stream{d_a, shared_w, 0, 0, n_blocks, 0}; };
  auto hbm = [&](uint8_t* b, int hint) { return Stream{b, slice, slice, 0, n_blocks, hint}; };
  struct Case { const char* name; Params p; };
  const Case cases[] = {
      {"read own (L2)", {own(0), none, none}},
      {"read shared weights (L2)", {shared(), none, none}},
      {"read hbm", {hbm(d_a, 0), none, none}},
      {"write own ring (L2)", {none, own(0), none}},
      {"write hbm", {none, hbm(d_b, 0), none}},
      {"read shared + write hbm (forward with stash)", {shared(), hbm(d_b, 0), none}},
      {"read hbm + write own ring", {hbm(d_a, 0), own(0), none}},
      {"read hbm + write own ring, hints", {hbm(d_a, 2), own(1), none}},
      {"read own + write own (all L2)", {own(0), own(0), none}},
      {"read hbm + write own ring + read nbr ring", {hbm(d_a, 0), own(0), nbr(0)}},
      {"read hbm + write own ring + read nbr ring, hints", {hbm(d_a, 2), own(1), nbr(1)}},
  };
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (const Case& c : cases) {
    float best = 1e30f;
    for (int it = 0; it < 3; ++it) {
      cudaEventRecord(e0);
      bw_kernel<<<n_ctas, 96, smem>>>(c.p);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); return 1; }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    const double per = (double)n_blocks * kBlk * n_ctas / best / 1e9;
    printf("{\"ctas\": %d, \"case\": \"%s\", \"ms\": %.3f, \"r1_TBps\": %.2f, \"w_TBps\": %.2f, \"r2_TBps\": %.2f, \"total_TBps\": %.2f}\n", n_ctas, c.name, best,
           c.p.r1.n ? per : 0.0, c.p.w.n ? per : 0.0, c.p.r2.n ? per : 0.0, per * ((c.p.r1.n > 0) + (c.p.w.n > 0) + (c.p.r2.n > 0)));
  }
  return 0;
}
