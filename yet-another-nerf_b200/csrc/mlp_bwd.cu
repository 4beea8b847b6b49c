// Backward of the fused NeRF-MLP for sm_100a (autograd of yanerf/pipelines/models/nerf_mlp.py:117-177).
//
// Three stages, all reading the 16-bit activation stash the forward kernel wrote:
//   1. mlp_bwd_dgrad_kernel   fused data-gradient chain, tile-pair persistent like the forward: colour head ->
//                             colour hidden -> intermediate -> trunk layers n-1..1.  A = dY tile in shared memory
//                             (K-major), B = transposed weight blocks streamed by TMA, fp32 accumulate in TMEM;
//                             the epilogue applies the ReLU mask (from the stash) and writes dY of the previous
//                             layer back to shared memory and to the gradient stash.
//   2. mlp_bwd_wgrad_kernel   per (layer, 128-output-feature half): dW = dY^T X over all points.  Both operands
//                             are the stashed [128 points x 64 features] blocks used as MN-major UMMA operands
//                             (the reduction runs over points), accumulators [128 x (256 + 64 + 16)] stay in TMEM
//                             across tiles; one fp32 atomic flush per CTA.  Bias gradients come for free from the
//                             constant-1 embedding channel (or a tile of ones for layers without it).
//   3. small SIMT kernels     the N=1 / N=3 heads and the per-ray direction part of LinearWithRepeat
//                             (models/utils.py:207-211).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "mlp_common.cuh"
#include "sm100_ptx.cuh"

namespace ynb {

// timing-experiment flags (YN_BWD_DEBUG) exist only in `make INSTRUMENT=1` builds: the product kernels carry no
// predicates for them
#ifdef YN_INSTRUMENT
__device__ __forceinline__ int debug_flags(const int flags) { return flags; }
#else
__device__ __forceinline__ int debug_flags(const int) { return 0; }
#endif

constexpr int kBwdRing = 4;
constexpr int kBwdThreads = 352;  // warp 0 TMA, warps 1 and 10 MMA issuers (tile 0 / 1), warps 2-9 epilogue
constexpr int kBSmemG = 0;                                  // [2][4][16 KB] dY tiles
constexpr int kBSmemRing = kBSmemG + 2 * 4 * kBlkBytes;     // [4][16 KB]
constexpr int kBSmemBar = kBSmemRing + kBwdRing * kBlkBytes;
constexpr int kBwdSmemBytes = kBSmemBar + 256 + 1024;

struct BwdParams {
  Arch arch;
  const float* directions;
  const float* rgb;
  const float* d_density;
  const float* d_rgb;
  const float* params;
  const uint8_t* wpack;
  const float* aux;
  const uint8_t* stash;
  uint8_t* gstash;
  float* grads;
  float* tbuf;  // [128][256] fp32: T = dY_colour^T relu(h_last) (see mlp_bwd_inter_kernel); then [128] column sums of dY_colour
  const float* amax;  // max |d_density|, |d_rgb| of this call (grad_amax_kernel); the fp16 gradient scale derives from it
  int64_t n_points;
  int64_t R;
  int P;
  int debug;  // YN_BWD_DEBUG (timing experiments, tools/bwd_split*.sh): 1/2/4/8 skip dgrad / wgrad / heads / direction; 16 heads without
              // prefetch; 32 / 64 gradient-stash writes / wgrad reads in an L2-resident window; 128 no masks; 256 no gradient-stash
              // stores; 512 wgrad without operand copies; 1024 wgrad without the small (embedding / ones) MMAs
};

__device__ __forceinline__ void bwd_named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bwd_st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ReLU masks come from the forward kernel's sign stash: bit (31 - j) of word q = "pre-activation of column 32 q + j was
// negative" (mlp_common.cuh).  Four columns at a time: the 4 bits are spread to the most significant bits of the 4 bytes
// of a word by one multiplication (bit i lands at 7 (i + 1) + i, no two products collide), and `prmt` in its
// sign-replicate mode turns a byte's top bit into a 0xFFFF / 0x0000 half-word mask.
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
// `p0` holds columns (j, j+1), `p1` columns (j+2, j+3) as packed 16-bit pairs, j a multiple of 4 inside mask word `w`
__device__ __forceinline__ void mask_quad(uint32_t w, int j, uint32_t& p0, uint32_t& p1) {
  const uint32_t t = (w >> (28 - j)) & 0xFu;        // bit 3 = column j ... bit 0 = column j + 3
  const uint32_t s = t * 0x10204080u;               // byte 3 msb = column j, byte 2 = j + 1, byte 1 = j + 2, byte 0 = j + 3
  p0 &= ~prmt(s, 0u, 0xAABBu);                      // low half <- sign(byte 3), high half <- sign(byte 2)
  p1 &= ~prmt(s, 0u, 0x8899u);                      // low half <- sign(byte 1), high half <- sign(byte 0)
}

// fp16 operands (kFmt 0) have 5 exponent bits: gradients d(loss)/d(activation) of a mean-reduced loss over thousands of rays
// (1e-6 .. 1e-9) would sit in or below the subnormal range.  Every 16-bit gradient of one backward call is therefore carried
// multiplied by ONE power of two S, chosen from the largest incoming gradient so that it lands at 2^-4 .. 2^-5 (20 binades
// of headroom for growth through the layers, 20 below before precision is lost), and every fp32 flush multiplies by 1/S:
// exact, since S is a power of two and all accumulation is fp32.  bf16 (kFmt 1) has fp32's exponent range: S = 1.
template <int kFmt>
__device__ __forceinline__ float grad_scale(const BwdParams& p) {
  if (kFmt == 1) return 1.f;
  const float a = *p.amax;
  if (!(a > 0.f) || !(a < 3.0e38f)) return 1.f;
  float e = -ceilf(log2f(a)) - 4.f;
  e = fminf(fmaxf(e, -100.f), 100.f);
  return exp2f(e);
}

__global__ void __launch_bounds__(256) grad_amax_kernel(const float* __restrict__ d_density, const float* __restrict__ d_rgb,
                                                        int64_t n_points, int C, float* amax) {
  float m = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_points; i += stride) {
    m = fmaxf(m, fabsf(__ldg(d_density + i)));
    for (int c = 0; c < C; ++c) m = fmaxf(m, fabsf(__ldg(d_rgb + i * C + c)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<unsigned int*>(amax), __float_as_uint(m));  // non-negative floats order like uints
}

// epilogue of one 128-column half of a data-gradient step.
// kMode 0: dY = acc; 1: dY = mask(acc); 2: dY = mask(acc + d_density * w_density)
template <int kFmt, int kMode>
__device__ __forceinline__ void dgrad_epilogue_half(uint32_t t_addr, int c_lo, const uint4 mask, uint32_t swz, float dd,
                                                    const float* __restrict__ wd, uint32_t g_row) {
  const uint32_t mw[4] = {mask.x, mask.y, mask.z, mask.w};
#pragma unroll
  for (int cb = 0; cb < 4; ++cb) {
    uint32_t v[32];
    tmem_ld32(t_addr + c_lo + cb * 32, v);
    tmem_ld_wait();
    const int c0 = c_lo + cb * 32;
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float x0 = __uint_as_float(v[j]), x1 = __uint_as_float(v[j + 1]);
      float x2 = __uint_as_float(v[j + 2]), x3 = __uint_as_float(v[j + 3]);
      if (kMode == 2) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(wd + c0 + j));
        x0 = fmaf(dd, w.x, x0); x1 = fmaf(dd, w.y, x1); x2 = fmaf(dd, w.z, x2); x3 = fmaf(dd, w.w, x3);
      }
      pk[j / 2] = Half2Pack<kFmt>::pack(x0, x1);
      pk[j / 2 + 1] = Half2Pack<kFmt>::pack(x2, x3);
      if (kMode != 0) mask_quad(mw[cb], j, pk[j / 2], pk[j / 2 + 1]);
    }
    const uint32_t blk = g_row + (c0 >> 6) * kBlkBytes;
    const uint32_t u0 = ((c0 >> 5) & 1) * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      bwd_st_shared_v4(blk + (((u0 + i) ^ swz) << 4), pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
  }
}

template <int kFmt>
__global__ void __launch_bounds__(kBwdThreads, 1) mlp_bwd_dgrad_kernel(const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_g = smem_base + kBSmemG;
  const uint32_t s_ring = smem_base + kBSmemRing;
  const uint32_t s_bar = smem_base + kBSmemBar;
  // barriers: full[4], empty[4], then per tile g: half_full[g][2], blk01_free[g], epi_done[g][2]
  const uint32_t bar_full = s_bar, bar_empty = s_bar + 8 * kBwdRing, bar_hfull = s_bar + 16 * kBwdRing,
                 bar_b01 = bar_hfull + 32, bar_epi = bar_b01 + 16, s_tmem_ptr = bar_epi + 32;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const Arch& A = p.arch;
  const int n = A.n_layers;
  const int n_steps = n + 1;  // step 0: merged colour hidden + intermediate layer; step 1: unused; steps 2..n: trunk n-1 .. 1
  const int64_t n_tiles = (p.n_points + kTileM - 1) / kTileM;
  const int64_t n_pairs = (n_tiles + 1) / 2;
  const int blocks_per_tile = A.stash_blocks_per_tile();

  if (threadIdx.x == 0) {
    for (int i = 0; i < kBwdRing; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 2);  // both MMA issuers release a weight block
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(bar_hfull + 16 * g, 1);
      mbar_init(bar_hfull + 16 * g + 8, 1);
      mbar_init(bar_b01 + 8 * g, 1);
      mbar_init(bar_epi + 16 * g, 128);
      mbar_init(bar_epi + 16 * g + 8, 128);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(s_tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(s_tmem_ptr));

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer: transposed weight blocks
    if (elect_one()) {
      uint32_t slot = 0, phase = 0;
      const uint8_t* wsrc = p.wpack + (size_t)A.total_stages() * kBlkBytes;
      int total = 0;
      for (int l = n + 1; l >= 1; --l) total += A.bwd_stages(l);
      for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        for (int s = 0; s < total; ++s) {
          mbar_wait(bar_empty + 8 * slot, phase ^ 1);
          mbar_arrive_expect_tx(bar_full + 8 * slot, kBlkBytes);
          bulk_g2s(s_ring + slot * kBlkBytes, wsrc + (size_t)s * kBlkBytes, kBlkBytes, bar_full + 8 * slot);
          if (++slot == kBwdRing) { slot = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == 10) {
    // ---------------------------------------------------------------- MMA issuers (one per tile of the pair)
    const int g = warp == 1 ? 0 : 1;
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc(128, 128, kFmt, 0, 0);
      const uint32_t my_epi = bar_epi + 16 * g, my_hfull = bar_hfull + 16 * g, my_b01 = bar_b01 + 8 * g;
      const uint32_t g_base = s_g + g * 4 * kBlkBytes;
      uint32_t slot = 0, phase = 0, ed_phase0 = 0, ed_phase1 = 0;
      for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        for (int st = 0; st < n_steps; ++st) {
          if (st == 1) continue;            // the intermediate layer is merged into the colour hidden layer (mlp_common.cuh)
          const int nkb = st == 0 ? 2 : 4;  // reduction over the layer's outputs (128 for the colour hidden layer)
          mbar_wait(my_epi, ed_phase0);
          ed_phase0 ^= 1;
          tc_fence_after();
          bool waited1 = false;
          for (int nh = 0; nh < 2; ++nh) {
            const uint32_t d_tmem = tmem_base + g * 256 + nh * 128;
            for (int kb = 0; kb < nkb; ++kb) {
              if (!waited1 && (nh == 1 || kb >= 2)) {
                mbar_wait(my_epi + 8, ed_phase1);
                ed_phase1 ^= 1;
                tc_fence_after();
                waited1 = true;
              }
              mbar_wait(bar_full + 8 * slot, phase);
              tc_fence_after();
              const uint64_t b_desc = umma_desc_kmajor(s_ring + slot * kBlkBytes);
              const uint64_t a_desc = umma_desc_kmajor(g_base + kb * kBlkBytes);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
              umma_commit(bar_empty + 8 * slot);
              if (++slot == kBwdRing) { slot = 0; phase ^= 1; }
              if (nh == 1 && kb == 1) umma_commit(my_b01);
            }
            umma_commit(my_hfull + 8 * nh);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 2 && warp <= 9) {
    // ---------------------------------------------------------------- epilogue groups
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;
    const uint32_t my_epi = bar_epi + 16 * g, my_hfull = bar_hfull + 16 * g, my_b01 = bar_b01 + 8 * g;
    const int row = q * 32 + lane;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * 256;
    const uint32_t g_g = s_g + g * 4 * kBlkBytes;
    const uint32_t swz = static_cast<uint32_t>(row & 7);
    const uint32_t g_row = g_g + static_cast<uint32_t>(row) * 128u;
    const bool leader = (warp - 2) % 4 == 0 && lane == 0;
    const float* wd = p.aux + A.aux_wd();
    const float* w2 = p.aux + A.aux_w2();
    const int C = A.color_dim;
    const float gscale = grad_scale<kFmt>(p);  // every 16-bit gradient of this call is carried times this power of two
    uint32_t hf_phase0 = 0, hf_phase1 = 0, b01_phase = 0;

    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      const int64_t tile = 2 * pair + g;
      const bool tile_real = tile < n_tiles;
      const bool tile_live = tile_real && !(debug_flags(p.debug) & 256);  // (bit 256, timing experiment: no gradient-stash stores)
      const int64_t gidx = tile * kTileM + row;
      const bool valid = gidx < p.n_points;
      // rows of partner tiles beyond the end alias tile 0 of the stash for reads; their gradients are zero
      const int64_t rtile = tile_real ? tile : 0;
      // this row's ReLU sign masks (32 B per layer): mask m at A.mask_offset(m)
      const uint8_t* mask_rows = p.stash + (size_t)rtile * blocks_per_tile * kBlkBytes + (size_t)row * 32;
      // (YN_BWD_DEBUG bit 32, timing experiment: every gradient-stash store lands in a 64-tile window that stays in L2)
      uint8_t* gstash_tile = p.gstash + (size_t)((debug_flags(p.debug) & 32) ? rtile % 64 : rtile) * blocks_per_tile * kBlkBytes;

      if (leader) bulk_wait_read<0>();
      bwd_named_bar_sync(1 + g, 128);
      // ---- head gradients: d(pre-sigmoid) -> d(colour hidden, pre-ReLU) = mask(hid) * (W2^T ds)
      float ds[4] = {0.f, 0.f, 0.f, 0.f};
      float dd = 0.f;
      if (valid) {
        dd = __ldg(p.d_density + gidx) * gscale;
        for (int c = 0; c < C; ++c) {
          const float y = __ldg(p.rgb + gidx * C + c);
          ds[c] = __ldg(p.d_rgb + gidx * C + c) * y * (1.f - y) * gscale;
        }
      }
      {
        const uint4 hm = __ldg(reinterpret_cast<const uint4*>(mask_rows + A.mask_offset(n)));  // colour hidden layer
#pragma unroll 1
        for (int u = 0; u < 16; ++u) {  // 16 units of 8 columns = 128 hidden features
          const int c0 = u * 8;
          const int q = u >> 2;
          const uint32_t mw = q == 0 ? hm.x : (q == 1 ? hm.y : (q == 2 ? hm.z : hm.w));
          float x[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] = 0.f;
          for (int c = 0; c < C; ++c) {
            const float4 wa = __ldg(reinterpret_cast<const float4*>(w2 + c * kDirPad + c0));
            const float4 wb = __ldg(reinterpret_cast<const float4*>(w2 + c * kDirPad + c0 + 4));
            x[0] = fmaf(ds[c], wa.x, x[0]); x[1] = fmaf(ds[c], wa.y, x[1]);
            x[2] = fmaf(ds[c], wa.z, x[2]); x[3] = fmaf(ds[c], wa.w, x[3]);
            x[4] = fmaf(ds[c], wb.x, x[4]); x[5] = fmaf(ds[c], wb.y, x[5]);
            x[6] = fmaf(ds[c], wb.z, x[6]); x[7] = fmaf(ds[c], wb.w, x[7]);
          }
          uint32_t p0 = Half2Pack<kFmt>::pack(x[0], x[1]), p1 = Half2Pack<kFmt>::pack(x[2], x[3]);
          uint32_t p2 = Half2Pack<kFmt>::pack(x[4], x[5]), p3 = Half2Pack<kFmt>::pack(x[6], x[7]);
          mask_quad(mw, c0 & 31, p0, p1);
          mask_quad(mw, (c0 & 31) + 4, p2, p3);
          bwd_st_shared_v4(g_row + (c0 >> 6) * kBlkBytes + ((((c0 >> 3) & 7) ^ swz) << 4), p0, p1, p2, p3);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(my_epi);
      mbar_arrive(my_epi + 8);
      bwd_named_bar_sync(1 + g, 128);
      if (leader && tile_live) {
        uint8_t* dst = gstash_tile + (size_t)A.stash_block_of_layer(n + 1) * kBlkBytes;
        bulk_s2g(dst, g_g, kBlkBytes);
        bulk_s2g(dst + kBlkBytes, g_g + kBlkBytes, kBlkBytes);
        bulk_commit();
      }

      for (int st = 0; st < n_steps; ++st) {
        if (st == 1) continue;                    // merged intermediate layer: no step of its own
        const int l = n + 1 - st;                 // layer whose data gradient was just multiplied
        // mma layer whose output gradient this epilogue produces; the colour step multiplies by W_ci^T = (W_c[:, :H] W_i)^T
        // and lands directly on the last trunk layer's output
        const int prev = st == 0 ? n - 1 : l - 1;
        const bool last = st == (n_steps == 2 ? 0 : n_steps - 1);  // (step 1 does not exist)
        // sign masks of both halves, requested before the wait for the accumulator
        const uint4* mp = reinterpret_cast<const uint4*>(mask_rows + A.mask_offset(prev));
        const uint4 mk0 = __ldg(mp), mk1 = __ldg(mp + 1);
        // ---- half 0
        mbar_wait(my_hfull, hf_phase0);
        hf_phase0 ^= 1;
        mbar_wait(my_b01, b01_phase);
        b01_phase ^= 1;
        tc_fence_after();
        // the gradient-stash store that last read dY blocks 0,1 has finished reading: the head store for step 0 (the only
        // group in flight), otherwise the previous step's blocks-0,1 group (its blocks-2,3 group, the most recent one, may
        // still be in flight)
        if (leader) {
          if (st == 0) bulk_wait_read<0>();
          else bulk_wait_read<1>();
        }
        bwd_named_bar_sync(1 + g, 128);
        // step 0 arrives at the last trunk layer's output: + the rank-1 density-head term
        if (debug_flags(p.debug) & 128) dgrad_epilogue_half<kFmt, 0>(t_row, 0, mk0, swz, dd, wd, g_row);  // (timing experiment: no masks)
        else if (st == 0) dgrad_epilogue_half<kFmt, 2>(t_row, 0, mk0, swz, dd, wd, g_row);
        else dgrad_epilogue_half<kFmt, 1>(t_row, 0, mk0, swz, dd, wd, g_row);
        tc_fence_before();
        fence_proxy_async_smem();
        if (!last) mbar_arrive(my_epi);
        // store blocks 0,1 right away: two 32 KB stores per step spread over time instead of one 64 KB burst
        bwd_named_bar_sync(1 + g, 128);
        uint8_t* gdst = gstash_tile + (size_t)A.stash_block_of_layer(prev) * kBlkBytes;
        if (leader) {
          if (tile_live) {
            bulk_s2g(gdst, g_g, kBlkBytes);
            bulk_s2g(gdst + kBlkBytes, g_g + kBlkBytes, kBlkBytes);
          }
          bulk_commit();
        }
        // ---- half 1
        mbar_wait(my_hfull + 8, hf_phase1);
        hf_phase1 ^= 1;
        tc_fence_after();
        if (leader) bulk_wait_read<1>();  // the previous step's blocks-2,3 store (most recent group: this step's blocks 0,1)
        bwd_named_bar_sync(1 + g, 128);
        // step 0 arrives at the last trunk layer's output: + the rank-1 density-head term
        if (debug_flags(p.debug) & 128) dgrad_epilogue_half<kFmt, 0>(t_row, 128, mk1, swz, dd, wd, g_row);
        else if (st == 0) dgrad_epilogue_half<kFmt, 2>(t_row, 128, mk1, swz, dd, wd, g_row);
        else dgrad_epilogue_half<kFmt, 1>(t_row, 128, mk1, swz, dd, wd, g_row);
        tc_fence_before();
        fence_proxy_async_smem();
        if (!last) mbar_arrive(my_epi + 8);
        bwd_named_bar_sync(1 + g, 128);
        if (leader) {
          if (tile_live) {
            bulk_s2g(gdst + 2 * kBlkBytes, g_g + 2 * kBlkBytes, kBlkBytes);
            bulk_s2g(gdst + 3 * kBlkBytes, g_g + 3 * kBlkBytes, kBlkBytes);
          }
          bulk_commit();
        }
      }
    }
    if (leader) bulk_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// weight gradient
// ------------------------------------------------------------------------------------------------
constexpr int kWgThreads = 192;
constexpr int kWgStageBytes = 7 * kBlkBytes;  // 2 dY blocks + 4 hidden blocks + 1 embedding block
constexpr int kWgStages = 2;
constexpr int kWgSmemOnes = kWgStages * kWgStageBytes;  // 1 KB of ones
constexpr int kWgSmemBar = kWgSmemOnes + 1024;
constexpr int kWgSmemBytes = kWgSmemBar + 128 + 1024;

// MN-major operand over [points x 64-feature blocks]: atoms of 64 features (128 B) x 8 points, 8-point groups 1024 B
// apart (SBO), 64-feature blocks 16 KB apart (LBO)
__device__ __forceinline__ uint64_t umma_desc_mnmajor(uint32_t smem_addr) { return umma_desc_sw128(smem_addr, kBlkBytes, 1024); }

template <int kFmt>
__global__ void __launch_bounds__(kWgThreads, 1) mlp_bwd_wgrad_kernel(const BwdParams p, int n_jobs, int n_splits) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_ones = smem_base + kWgSmemOnes;
  const uint32_t s_bar = smem_base + kWgSmemBar;
  const uint32_t bar_full = s_bar, bar_empty = s_bar + 16, bar_done = s_bar + 32, s_tmem_ptr = s_bar + 40;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const Arch& A = p.arch;
  const int n = A.n_layers;
  const int job = blockIdx.x % n_jobs;
  const int split = blockIdx.x / n_jobs;
  // job -> (layer, output half) for the n trunk layers, then the colour hidden layer (a single half).  The intermediate
  // layer has no job: its gradient comes out of the colour job's product (mlp_bwd_inter_kernel).
  const int l = job < 2 * n ? job / 2 : n + 1;
  const int mh = job < 2 * n ? job % 2 : 0;
  const bool color_job = l == n + 1;
  const bool has_hidden = l >= 1;
  const bool has_emb = A.has_emb(l);
  const bool use_ones = !has_emb;  // bias gradient via a tile of ones when there is no constant-1 channel
  const int64_t n_tiles = (p.n_points + kTileM - 1) / kTileM;
  const int blocks_per_tile = A.stash_blocks_per_tile();
  // X_l: output of the previous mma layer; the colour job multiplies against the LAST TRUNK layer's output instead of the
  // intermediate output (which is not stashed): T = dY_c^T relu(h_last)
  const int xprev = color_job ? n - 1 : l - 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kWgStages; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    mbar_init(bar_done, 1);
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {  // 1 KB of 16-bit ones
    const uint32_t one2 = Half2Pack<kFmt>::pack(1.f, 1.f);
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(s_ones + 4 * i), "r"(one2) : "memory");
  }
  fence_proxy_async_smem();
  if (warp == 1) {
    tmem_alloc(s_tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(s_tmem_ptr));

  const uint32_t stage_bytes = (2 + (has_hidden ? 4 : 0) + (has_emb ? 1 : 0)) * kBlkBytes;
  if (warp == 0) {
    if (elect_one()) {
      uint32_t slot = 0, phase = 0;
      for (int64_t t = split; t < n_tiles; t += n_splits) {
        // (YN_BWD_DEBUG bit 64, timing experiment: operands come from a 64-tile window that stays in L2)
        const int64_t ts = (debug_flags(p.debug) & 64) ? t % 64 : t;
        const uint8_t* st = p.stash + (size_t)ts * blocks_per_tile * kBlkBytes;
        const uint8_t* gs = p.gstash + (size_t)ts * blocks_per_tile * kBlkBytes;
        const uint32_t dst = smem_base + slot * kWgStageBytes;
        const uint32_t bar = bar_full + 8 * slot;
        mbar_wait(bar_empty + 8 * slot, phase ^ 1);
        if (debug_flags(p.debug) & 512) {  // (timing experiment: no operand copies, the MMA pipeline alone)
          mbar_arrive(bar);
          if (++slot == kWgStages) { slot = 0; phase ^= 1; }
          continue;
        }
        mbar_arrive_expect_tx(bar, stage_bytes);
        bulk_g2s(dst, gs + (size_t)(A.stash_block_of_layer(l) + 2 * mh) * kBlkBytes, 2 * kBlkBytes, bar);
        if (has_hidden) bulk_g2s(dst + 2 * kBlkBytes, st + (size_t)A.stash_block_of_layer(xprev) * kBlkBytes, 4 * kBlkBytes, bar);
        if (has_emb) bulk_g2s(dst + 6 * kBlkBytes, st, kBlkBytes, bar);
        if (++slot == kWgStages) { slot = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc256 = umma_idesc(128, 256, kFmt, 1, 1);
      constexpr uint32_t idesc64 = umma_idesc(128, 64, kFmt, 1, 1);
      constexpr uint32_t idesc16 = umma_idesc(128, 16, kFmt, 1, 1);
      uint32_t slot = 0, phase = 0;
      uint32_t first = 1;
      const uint64_t ones_desc = umma_desc_sw128(s_ones, 128, 256) & ~(static_cast<uint64_t>(7) << 61);  // no swizzle
      for (int64_t t = split; t < n_tiles; t += n_splits) {
        const uint32_t base = smem_base + slot * kWgStageBytes;
        mbar_wait(bar_full + 8 * slot, phase);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // 16 points per MMA
          const uint64_t a_desc = umma_desc_mnmajor(base + k * 2048);
          const uint32_t acc = (first && k == 0) ? 0u : 1u;
          if (has_hidden) umma_f16(tmem_base, a_desc, umma_desc_mnmajor(base + 2 * kBlkBytes + k * 2048), idesc256, acc);
          if (has_emb && !(debug_flags(p.debug) & 1024)) umma_f16(tmem_base + 256, a_desc, umma_desc_mnmajor(base + 6 * kBlkBytes + k * 2048), idesc64, acc);
          if (use_ones && !(debug_flags(p.debug) & 1024)) umma_f16(tmem_base + 320, a_desc, ones_desc, idesc16, acc);  // (1024: timing experiment)
        }
        first = 0;
        umma_commit(bar_empty + 8 * slot);
        if (++slot == kWgStages) { slot = 0; phase ^= 1; }
      }
      umma_commit(bar_done);
    }
    __syncwarp();
  } else if (split < n_tiles) {
    // ---------------------------------------------------------------- flush: TMEM -> fp32 atomics
    const int q = warp & 3;
    const int out = mh * 128 + q * 32 + lane;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    mbar_wait(bar_done, 0);
    tc_fence_after();
    const float inv = 1.f / grad_scale<kFmt>(p);  // the accumulators hold S * gradient (exact un-scaling: S is a power of two)
    const int dout = A.dout(l), din = A.din(l), hin = A.hidden_in(l), exyz = A.embed_xyz();
    float* W = color_job ? p.tbuf + (int64_t)out * kInner : p.grads + A.w_offset(l) + (int64_t)out * din;
    float* bgrad = p.grads + A.b_offset(l) + out;
    const bool row_ok = out < dout;
    if (has_hidden) {
      for (int cb = 0; cb < 8; ++cb) {
        uint32_t v[32];
        tmem_ld32(t_row + cb * 32, v);
        tmem_ld_wait();
        if (row_ok)
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (cb * 32 + j < hin) atomicAdd(W + cb * 32 + j, __uint_as_float(v[j]) * inv);
      }
    }
    if (has_emb) {
      for (int cb = 0; cb < 2; ++cb) {
        uint32_t v[32];
        tmem_ld32(t_row + 256 + cb * 32, v);
        tmem_ld_wait();
        if (row_ok)
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int e = cb * 32 + j;
            if (e < exyz) atomicAdd(W + hin + e, __uint_as_float(v[j]) * inv);
            else if (e == 63) atomicAdd(bgrad, __uint_as_float(v[j]) * inv);  // constant-1 channel -> bias
          }
      }
    }
    if (use_ones) {
      uint32_t v[32];
      tmem_ld32(t_row + 320, v);  // columns 320..335 hold 16 identical sums (the rest is unused)
      tmem_ld_wait();
      if (row_ok) {
        atomicAdd(bgrad, __uint_as_float(v[0]) * inv);
        if (color_job) atomicAdd(p.tbuf + kDirPad * kInner + out, __uint_as_float(v[0]) * inv);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// small heads: density layer (N = 1), colour output layer (N = color_dim)
// ------------------------------------------------------------------------------------------------
template <int kFmt>
__device__ __forceinline__ float half_at(const uint8_t* tile_layer, int row, int col) {
  const uint16_t bits = *reinterpret_cast<const uint16_t*>(tile_layer + (size_t)(col >> 6) * kBlkBytes + sw128_offset(row, col & 63));
  if (kFmt == 1) return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&bits));
  return __half2float(*reinterpret_cast<const __half*>(&bits));
}

template <int kFmt>
__device__ __forceinline__ float half_bits_to_float(uint16_t bits) {
  if (kFmt == 1) return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&bits));
  return __half2float(*reinterpret_cast<const __half*>(&bits));
}

// Density head (N = 1) and colour output layer (N = color_dim) weight gradients:
//   dwd[j] += sum_r dd[r] * feat[r][j],  dW2[c][j] += sum_r ds[r][c] * hid[r][j]   (+ the two bias gradients)
// Thread (rg, u) of a 256-thread block owns the u-th 16-byte unit (8 columns) of the 512-byte activation row and the
// rows rg, rg+8, ... of every tile: a warp reads whole rows (4 x 128-byte lines), all loads of a tile are in flight
// together; the 8 row groups are folded through shared memory once, at the end.
template <int kFmt>
__global__ void __launch_bounds__(256, 2) mlp_bwd_heads_kernel(const BwdParams p) {
  __shared__ float s_dd[kTileM];
  __shared__ float s_ds[kTileM][4];
  __shared__ float s_red[8][kInner + 3 * kDirPad];
  const Arch& A = p.arch;
  const int n = A.n_layers, C = A.color_dim;
  const int64_t n_tiles = (p.n_points + kTileM - 1) / kTileM;
  const int blocks_per_tile = A.stash_blocks_per_tile();
  const int t = threadIdx.x;
  const int rg = t >> 5, u = t & 31;          // row group, 16-byte unit of the row (columns 8u .. 8u+7)
  const size_t unit_off = (size_t)(u >> 3) * kBlkBytes;
  const uint32_t u_in_blk = u & 7;
  const bool do_hid = u < kDirPad / 8;
  float acc_wd[8], acc_w2[3][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc_wd[i] = 0.f;
    acc_w2[0][i] = acc_w2[1][i] = acc_w2[2][i] = 0.f;
  }
  float acc_bd = 0.f, acc_b2 = 0.f;
  // d(density) and d(pre-sigmoid colour) of one point of a tile; loaded one tile ahead so that this (dependent, three
  // small arrays) global-load latency overlaps the previous tile's work
  float nx_dd = 0.f, nx_ds[3] = {0.f, 0.f, 0.f};
  auto fetch = [&](int64_t tile) {
    nx_dd = 0.f;
    nx_ds[0] = nx_ds[1] = nx_ds[2] = 0.f;
    const int64_t gidx = tile * kTileM + t;
    if (t < kTileM && tile < n_tiles && gidx < p.n_points) {
      nx_dd = __ldg(p.d_density + gidx);
      for (int c = 0; c < C; ++c) {
        const float y = __ldg(p.rgb + gidx * C + c);
        nx_ds[c] = __ldg(p.d_rgb + gidx * C + c) * y * (1.f - y);
      }
    }
  };
  fetch(blockIdx.x);
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();
    if (debug_flags(p.debug) & 16) fetch(tile);  // experiment: no prefetch
    if (t < kTileM) {
      s_dd[t] = nx_dd;
      s_ds[t][0] = nx_ds[0]; s_ds[t][1] = nx_ds[1]; s_ds[t][2] = nx_ds[2]; s_ds[t][3] = 0.f;
    }
    fetch(tile + gridDim.x);
    __syncthreads();
    const uint8_t* base = p.stash + (size_t)tile * blocks_per_tile * kBlkBytes;
    const uint8_t* feat = base + (size_t)A.stash_block_of_layer(n - 1) * kBlkBytes + unit_off;  // last trunk output
    const uint8_t* hid = base + (size_t)A.stash_block_of_layer(n + 1) * kBlkBytes + unit_off;   // colour hidden
    uint4 f[16], h[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const uint32_t r = rg + 8 * i;
      const size_t off = (size_t)r * 128 + ((u_in_blk ^ (r & 7)) << 4);
      f[i] = __ldg(reinterpret_cast<const uint4*>(feat + off));
      h[i] = do_hid ? __ldg(reinterpret_cast<const uint4*>(hid + off)) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int r = rg + 8 * i;
      const float dd = s_dd[r];
      const float d0 = s_ds[r][0], d1 = s_ds[r][1], d2 = s_ds[r][2];
      const uint32_t fw[4] = {f[i].x, f[i].y, f[i].z, f[i].w}, hw[4] = {h[i].x, h[i].y, h[i].z, h[i].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 fv = Half2Pack<kFmt>::unpack(fw[k]), hv = Half2Pack<kFmt>::unpack(hw[k]);
        acc_wd[2 * k] = fmaf(dd, fv.x, acc_wd[2 * k]);
        acc_wd[2 * k + 1] = fmaf(dd, fv.y, acc_wd[2 * k + 1]);
        acc_w2[0][2 * k] = fmaf(d0, hv.x, acc_w2[0][2 * k]); acc_w2[0][2 * k + 1] = fmaf(d0, hv.y, acc_w2[0][2 * k + 1]);
        acc_w2[1][2 * k] = fmaf(d1, hv.x, acc_w2[1][2 * k]); acc_w2[1][2 * k + 1] = fmaf(d1, hv.y, acc_w2[1][2 * k + 1]);
        acc_w2[2][2 * k] = fmaf(d2, hv.x, acc_w2[2][2 * k]); acc_w2[2][2 * k + 1] = fmaf(d2, hv.y, acc_w2[2][2 * k + 1]);
      }
    }
    if (t == 0)
      for (int r = 0; r < kTileM; ++r) acc_bd += s_dd[r];
    if (t >= 1 && t <= C)
      for (int r = 0; r < kTileM; ++r) acc_b2 += s_ds[r][t - 1];
  }
  // fold the 8 row groups
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s_red[rg][u * 8 + i] = acc_wd[i];
    if (do_hid)
#pragma unroll
      for (int c = 0; c < 3; ++c) s_red[rg][kInner + c * kDirPad + u * 8 + i] = acc_w2[c][i];
  }
  __syncthreads();
  // every block adds to the same 640 addresses at about the same time: start each block at a different offset so that
  // concurrent atomics land on different addresses (same-address atomics serialise in L2)
  constexpr int kFlush = kInner + 3 * kDirPad;
  const int rot = (int)((blockIdx.x * 83u) % kFlush);
  for (int j0 = t; j0 < kFlush; j0 += blockDim.x) {
    const int j = j0 + rot < kFlush ? j0 + rot : j0 + rot - kFlush;
    float v = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) v += s_red[g][j];
    if (j < kInner) {
      if (j < A.hidden_last) atomicAdd(p.grads + A.density_w_offset() + j, v);
    } else {
      const int c = (j - kInner) / kDirPad, jj = (j - kInner) % kDirPad;
      if (c < C && jj < A.hidden_dir) atomicAdd(p.grads + A.color2_w_offset() + (int64_t)c * A.hidden_dir + jj, v);
    }
  }
  if (t == 0) atomicAdd(p.grads + A.density_b_offset(), acc_bd);
  if (t >= 1 && t <= C) atomicAdd(p.grads + A.color2_b_offset() + t - 1, acc_b2);
}

// per-ray direction part of the colour hidden layer: dW_c[:, H + k] += sum_rays (sum_samples dY[ray, s, :]) emb27[ray][k]
// one warp per ray; lane owns 4 of the 128 hidden columns (one 8-byte load per sample, 16 samples in flight).  The outer
// product with the ray's 27 embedding channels is accumulated in REGISTERS over all rays of the warp (108 accumulators
// per lane); the 8 warps of a block are folded through shared memory once at the end.  (Per-ray shared-memory atomics, the
// first version, were the bottleneck: 3 456 of them per ray.)
template <int kFmt>
__global__ void __launch_bounds__(256) mlp_bwd_dir_kernel(const BwdParams p) {
  constexpr int kE = 28;                 // 27 direction-embedding channels (+pad); check_arch: embed_dir <= 32, here <= 27
  __shared__ float s_acc[kDirPad * kE];  // [j][k]
  const Arch& A = p.arch;
  const int n = A.n_layers;
  const int ed = A.embed_dir() < kE ? A.embed_dir() : kE;
  const int blocks_per_tile = A.stash_blocks_per_tile();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = threadIdx.x; i < kDirPad * kE; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const size_t layer_off = (size_t)A.stash_block_of_layer(n + 1) * kBlkBytes;
  const int c0 = lane * 4;  // columns c0..c0+3 live in one 16-byte unit
  const size_t col_off = (size_t)(c0 >> 6) * kBlkBytes + ((c0 & 7) << 1);
  const uint32_t unit = (c0 >> 3) & 7;
  float acc[4][kE];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < kE; ++k) acc[i][k] = 0.f;
  for (int64_t ray = (int64_t)blockIdx.x * nw + wib; ray < p.R; ray += (int64_t)gridDim.x * nw) {
    float gsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int s0 = 0; s0 < p.P; s0 += 16) {
      uint2 v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int64_t gidx = ray * p.P + min(s0 + i, p.P - 1);
        const int64_t t = gidx / kTileM;
        const uint32_t r = (uint32_t)(gidx % kTileM);
        v[i] = __ldg(reinterpret_cast<const uint2*>(p.gstash + (size_t)t * blocks_per_tile * kBlkBytes + layer_off + col_off +
                                                    (size_t)r * 128 + ((unit ^ (r & 7)) << 4)));
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (s0 + i < p.P) {
          const float2 a = Half2Pack<kFmt>::unpack(v[i].x), b = Half2Pack<kFmt>::unpack(v[i].y);
          gsum[0] += a.x; gsum[1] += a.y; gsum[2] += b.x; gsum[3] += b.y;
        }
      }
    }
    // direction embedding of this ray (same arithmetic as dirbias_kernel)
    const float dx = p.directions[ray * 3], dy = p.directions[ray * 3 + 1], dz = p.directions[ray * 3 + 2];
    const float nrm = fmaxf(sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz))), 1e-12f);
    const float d[3] = {dx / nrm, dy / nrm, dz / nrm};
    const int nf = A.n_freq_dir;
    // lane k (< 27) evaluates embedding channel k once; every lane then reads all of them by shuffle
    float e_mine = 0.f;
    if (lane < ed) {
      const int k = lane;
      if (k < 3 * nf) e_mine = sinf(d[k / nf] * exp2f((float)(k % nf)));
      else if (k < 6 * nf) e_mine = cosf(d[(k - 3 * nf) / nf] * exp2f((float)((k - 3 * nf) % nf)));
      else e_mine = d[k - 6 * nf];
    }
#pragma unroll
    for (int k = 0; k < kE; ++k) {
      const float e = __shfl_sync(0xffffffffu, e_mine, k);  // 0 for k >= ed
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i][k] = fmaf(gsum[i], e, acc[i][k]);
    }
  }
  const float inv = 1.f / grad_scale<kFmt>(p);  // the gradient stash holds S * dY
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < kE; ++k) atomicAdd(&s_acc[(c0 + i) * kE + k], acc[i][k] * inv);
  __syncthreads();
  const int din = A.din(n + 1);
  float* W = p.grads + A.w_offset(n + 1);
  const int n_flush = A.hidden_dir * ed;
  const int rot = (int)((blockIdx.x * 331u) % (unsigned)n_flush);  // de-correlate the blocks' atomics (see heads kernel)
  for (int i0 = threadIdx.x; i0 < n_flush; i0 += blockDim.x) {
    const int i = i0 + rot < n_flush ? i0 + rot : i0 + rot - n_flush;
    const int j = i / ed, k = i % ed;
    atomicAdd(W + (int64_t)j * din + A.hidden_last + k, s_acc[j * kE + k]);
  }
}

// The intermediate layer is linear (no activation) and feeds only the colour hidden layer:
//   inter = relu(h) W_i^T + b_i,   d_inter = dY_c W_c[:, :H]        (h = last trunk output, H = hidden_last)
// so with T = dY_c^T relu(h)  [hidden_dir x H]  and  s = sum_points dY_c  [hidden_dir]  (both accumulated by the colour
// weight-gradient job over the trunk output the intermediate layer's own job would have read):
//   dW_c[:, :H] = dY_c^T inter = T W_i^T + s b_i^T,    dW_i = d_inter^T relu(h) = W_c[:, :H]^T T,    db_i = W_c[:, :H]^T s.
// Neither the intermediate output nor its gradient has to be written to HBM (16 blocks of 16 KB per tile less traffic),
// and the two products no longer see a 16-bit rounding of inter / d_inter.  fp32 master weights, two 128x256x256
// contractions: microseconds.  Block b < H: row b of dW_i and db_i[b]; block H + j: row j of dW_c[:, :H].
__global__ void __launch_bounds__(256) mlp_bwd_inter_kernel(const BwdParams p) {
  const Arch& A = p.arch;
  const int n = A.n_layers, H = A.hidden_last, D = A.hidden_dir;
  const int dinc = A.din(n + 1);
  const float* Wi = p.params + A.w_offset(n);
  const float* bi = p.params + A.b_offset(n);
  const float* Wc = p.params + A.w_offset(n + 1);
  const float* T = p.tbuf;
  const float* svec = p.tbuf + kDirPad * kInner;
  __shared__ float s_row[kInner];
  const int t = threadIdx.x;
  if ((int)blockIdx.x < H) {
    const int o = blockIdx.x;
    if (t < D) s_row[t] = Wc[(int64_t)t * dinc + o];  // column o of W_c
    __syncthreads();
    if (t < H) {
      float acc = 0.f;
      for (int j = 0; j < D; ++j) acc = fmaf(s_row[j], T[j * kInner + t], acc);
      p.grads[A.w_offset(n) + (int64_t)o * H + t] += acc;
    }
    if (t == 0) {
      float acc = 0.f;
      for (int j = 0; j < D; ++j) acc = fmaf(s_row[j], svec[j], acc);
      p.grads[A.b_offset(n) + o] += acc;
    }
  } else {
    // row j of dW_c[:, :H] = sum_k T[j][k] W_i[o][k] + s[j] b_i[o]: the contraction runs along the fast axis of W_i, so W_i
    // goes through shared memory in [256 o x 32 k] tiles (coalesced 128-byte row reads, conflict-free column reads)
    const int j = blockIdx.x - H;
    __shared__ float s_w[kInner][33];
    s_row[t] = t < H ? T[j * kInner + t] : 0.f;
    float acc = t < H ? svec[j] * bi[t] : 0.f;
    const int w = t >> 5, l = t & 31;
    for (int k0 = 0; k0 < H; k0 += 32) {
      __syncthreads();
#pragma unroll 8
      for (int r = 0; r < 32; ++r) {
        const int o = w * 32 + r;
        s_w[o][l] = (o < H && k0 + l < H) ? Wi[(int64_t)o * H + k0 + l] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < 32; ++kk) acc = fmaf(s_row[k0 + kk], s_w[t][kk], acc);
    }
    if (t < H) p.grads[A.w_offset(n + 1) + (int64_t)j * dinc + t] += acc;
  }
}

constexpr size_t kTbufBytes = (size_t)(kDirPad * kInner + kDirPad + 4) * sizeof(float);  // T, column sums, then the gradient amax word

static int launch_bwd(const BwdParams& p, cudaStream_t stream) {
  const Arch& A = p.arch;
  const int64_t n_tiles = (p.n_points + kTileM - 1) / kTileM;
  const int64_t n_pairs = (n_tiles + 1) / 2;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)(n_pairs < sms ? n_pairs : sms);
  const int n_jobs = 2 * A.n_layers + 1;
  int n_splits = sms / n_jobs;
  if (n_splits < 1) n_splits = 1;
  if (n_splits > n_tiles) n_splits = (int)n_tiles;
  auto run = [&](auto dgrad, auto wgrad, auto heads, auto dir) {
    if (first_use(reinterpret_cast<const void*>(dgrad)))
      cudaFuncSetAttribute(dgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmemBytes);
    if (first_use(reinterpret_cast<const void*>(wgrad)))
      cudaFuncSetAttribute(wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBytes);
    // YN_BWD_DEBUG (timing experiments only, wrong gradients; tools/bwd_split.sh): bit mask of kernels to skip
    const int skip = p.debug;
    cudaMemsetAsync(p.tbuf, 0, kTbufBytes, stream);
    if (A.fmt == 0)  // fp16 operands: find the scale of this call's 16-bit gradients (see grad_scale)
      grad_amax_kernel<<<2 * sms, 256, 0, stream>>>(p.d_density, p.d_rgb, p.n_points, A.color_dim, const_cast<float*>(p.amax));
    if (!(skip & 1)) dgrad<<<grid, kBwdThreads, kBwdSmemBytes, stream>>>(p);
    if (!(skip & 2)) {
      wgrad<<<n_jobs * n_splits, kWgThreads, kWgSmemBytes, stream>>>(p, n_jobs, n_splits);
      mlp_bwd_inter_kernel<<<A.hidden_last + A.hidden_dir, 256, 0, stream>>>(p);
    }
    // exactly one resident wave of head blocks (a partial second wave would run at a third of the occupancy)
    static int heads_per_sm = 0;
    if (heads_per_sm == 0) {
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&heads_per_sm, heads, 256, 0);
      if (heads_per_sm < 1) heads_per_sm = 1;
    }
    const int64_t heads_grid = (int64_t)heads_per_sm * sms;
    if (!(skip & 4)) heads<<<(int)(n_tiles < heads_grid ? n_tiles : heads_grid), 256, 0, stream>>>(p);
    const int64_t ray_blocks = (p.R + 7) / 8;
    if (!(skip & 8)) dir<<<(int)(ray_blocks < sms ? ray_blocks : sms), 256, 0, stream>>>(p);  // 254 registers: one block per SM
  };
  if (A.fmt == 1)
    run(mlp_bwd_dgrad_kernel<1>, mlp_bwd_wgrad_kernel<1>, mlp_bwd_heads_kernel<1>, mlp_bwd_dir_kernel<1>);
  else
    run(mlp_bwd_dgrad_kernel<0>, mlp_bwd_wgrad_kernel<0>, mlp_bwd_heads_kernel<0>, mlp_bwd_dir_kernel<0>);
  return check_launch("yn_mlp_bwd");
}

}  // namespace ynb

extern "C" int64_t yn_mlp_bwd_workspace_bytes(const yn_mlp_arch* arch, int64_t n_points) {
  // the gradient stash mirrors the activation stash (one 16-bit dY image per layer and tile); then the fp32 scratch of
  // mlp_bwd_inter_kernel
  const int64_t sb = yn_mlp_stash_bytes(arch, n_points);
  return sb < 0 ? sb : sb + (int64_t)ynb::kTbufBytes;
}

extern "C" int yn_mlp_bwd(const yn_mlp_arch* arch, const float* directions, const float* rgb, const float* d_density,
                          const float* d_rgb, const float* params, const void* wpack, const float* aux,
                          const void* stash, void* workspace, float* grads, int64_t R, int P, void* stream) {
  if (int rc = ynb::check_arch(arch)) return rc;
  if (R < 0 || P <= 0) return ynb::fail(YN_ERR_INVALID_ARGUMENT, "yn_mlp_bwd: bad sizes R=%lld P=%d", (long long)R, P);
  if (R == 0) return YN_OK;
  if (!directions || !rgb || !d_density || !d_rgb || !params || !wpack || !aux || !stash || !workspace || !grads)
    return ynb::fail(YN_ERR_INVALID_ARGUMENT, "yn_mlp_bwd: null pointer");
  ynb::BwdParams p;
  p.arch = ynb::arch_from_c(arch);
  p.directions = directions;
  p.rgb = rgb;
  p.d_density = d_density;
  p.d_rgb = d_rgb;
  p.params = params;
  p.wpack = static_cast<const uint8_t*>(wpack);
  p.aux = aux;
  p.stash = static_cast<const uint8_t*>(stash);
  p.gstash = static_cast<uint8_t*>(workspace);
  p.tbuf = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + yn_mlp_stash_bytes(arch, R * P));
  p.amax = p.tbuf + ynb::kDirPad * ynb::kInner + ynb::kDirPad;
  p.grads = grads;
  p.n_points = R * P;
  p.R = R;
  p.P = P;
  static const int debug = getenv("YN_BWD_DEBUG") ? atoi(getenv("YN_BWD_DEBUG")) : 0;
  p.debug = debug;
#ifndef YN_INSTRUMENT
  static bool warned = false;
  if ((debug & ~15) && !warned) {  // (bits 1/2/4/8 skip whole launches on the host and work in every build)
    warned = true;
    fprintf(stderr, "yn_mlp_bwd: YN_BWD_DEBUG bits above 8 need a `make INSTRUMENT=1` build of the library; ignored\n");
  }
#endif
  return ynb::launch_bwd(p, static_cast<cudaStream_t>(stream));
}
