// Fused NeRF-MLP forward for sm_100a.
//
// Replaces NeRFMLP.forward (yanerf/pipelines/models/nerf_mlp.py:117-177) together with
// ray_bundle_to_ray_points and HarmonicEmbedding.forward (models/utils.py:90-103,214-245):
// points = o + z*d, 63-channel harmonic embedding, the 8x256 skip trunk, density head, intermediate
// linear, LinearWithRepeat colour hidden layer and the sigmoid colour head, in ONE persistent kernel.
//
// One CTA per SM, 352 threads:
//   warp 0      TMA producer: streams 16 KB weight blocks [128 n x 64 k] (pre-swizzled image) through a
//               4-deep shared-memory ring with cp.async.bulk + mbarrier complete_tx.
//   warps 1,10  MMA issuers, one per tile: an elected thread issues tcgen05.mma (M=128, N=128, K=16, fp32 accumulate
//               in TMEM); A = the tile's activation blocks in shared memory (K-major, 128B swizzle), B = ring slot.
//   warps 2-5   epilogue group 0, warps 6-9 epilogue group 1: tcgen05.ld the accumulator, add bias, ReLU,
//               convert to 16 bit and write the next layer's A operand back to shared memory.  The first
//               "epilogue" of a tile computes the embedding, the last ones compute the fp32 heads.
// Two 128-point tiles are in flight per CTA (TMEM columns [0,256) and [256,512)) and advance in lockstep: every
// weight block read from L2 feeds the MMAs of both tiles.  Each 256-wide layer is issued as two 128-column halves;
// the epilogue of half 0 (-> activation blocks 0,1 of the next layer) runs while half 1 is still being multiplied,
// and the next layer starts on blocks 0,1 while the epilogue of half 1 fills blocks 2,3.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "mlp_common.cuh"
#include "sm100_ptx.cuh"

namespace ynb {

constexpr int kRing = 4;
constexpr int kFwdThreads = 352;  // warp 0 TMA, warps 1 and 10 MMA issuers (tile 0 / tile 1), warps 2-9 epilogue
constexpr int kSmemAct = 0;                              // [2][4][16 KB]
constexpr int kSmemEmb = kSmemAct + 2 * 4 * kBlkBytes;   // [2][16 KB]
constexpr int kSmemRing = kSmemEmb + 2 * kBlkBytes;      // [4][16 KB]
constexpr int kSmemBar = kSmemRing + kRing * kBlkBytes;  // barriers
constexpr int kSmemTotal = kSmemBar + 256;
constexpr int kFwdSmemBytes = kSmemTotal + 1024;         // + alignment slack

struct FwdParams {
  Arch arch;
  const float* origins;
  const float* directions;
  const float* lengths;
  const float* dirbias;  // [R][128]
  const uint8_t* wpack;
  const float* aux;
  float* density;
  float* rgb;
  uint8_t* stash;  // nullptr in inference
  int64_t n_points;
  int P;
  long long* trace;  // timing experiments only (YN_FWD_TRACE=<file>): CTA 0 logs (tag, clock64) pairs, 4 roles x 2048 events
  int debug;  // timing experiments only (YN_FWD_DEBUG bit mask): 1 = epilogue skips TMEM/STS work, 2 = producer skips the copies, 4 = stash into an L2-resident window
};

// The timing experiments (YN_FWD_TRACE timeline, YN_FWD_DEBUG work-skipping flags) are compiled in only with
// -DYN_INSTRUMENT (`make INSTRUMENT=1`): in the product build their predicates and skipped blocks sat in the hot loops
// of the issuers and the epilogue warps and cost 1.8 % of the render (interleaved A/B on one box, tools/ab_bench.sh).
#ifdef YN_INSTRUMENT
constexpr bool kInstrument = true;
#else
constexpr bool kInstrument = false;
#endif
constexpr int kTraceEvents = 2048;
struct Tracer {
  long long* buf;
  int n;
  __device__ __forceinline__ void init(long long* base, int role) {
    if (!kInstrument) return;
    buf = (base && blockIdx.x == 0) ? base + (size_t)role * 2 * kTraceEvents : nullptr;
    n = 0;
  }
  __device__ __forceinline__ void log(int tag) {
    if (!kInstrument) return;
    if (buf && n < kTraceEvents) {
      buf[2 * n] = tag;
      buf[2 * n + 1] = clock64();
      ++n;
    }
  }
};
__device__ __forceinline__ int debug_flags(const int flags) { return kInstrument ? flags : 0; }
// YN_FWD_DEBUG bit 128 (with YN_FWD_TRACE): no event log; CTA 0 sums the cycles its issuer 0 / epilogue group 0 / producer
// spend in each kind of barrier wait and writes the sums to the head of their trace rows (tools/fwd_waits.py)
#define YN_TIMED(on, acc, stmt)                 \
  do {                                          \
    long long t0_ = (on) ? clock64() : 0;       \
    stmt;                                       \
    if (on) (acc) += clock64() - t0_;           \
  } while (0)

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// sin / cos of x * 2^k for the harmonic embedding.  The reference evaluates sin(fl32(x * 2^k)) and x * 2^k is exact, so
// the phase is reduced once per coordinate: x / (2 pi) as an unevaluated fp32 pair (hi + lo, ~48 bits; this GPU's fp64
// rate is too low for the job), scaled by 2^k exactly, fractional part of `hi` exact, `lo` added back.  Remaining error:
// one fp32 rounding of the reduced angle in [-pi, pi] (1.9e-7) + the SFU approximation (4e-7 abs on that range):
// < 1e-6 absolute at every octave, 1/250 of the fp16 operand ulp.  ~8 instructions instead of sincosf's ~50.
constexpr float kInvTwoPiHi = 0.15915494309189535f;
constexpr float kInvTwoPiLo = static_cast<float>(0.15915494309189533577 - static_cast<double>(kInvTwoPiHi));
struct Turns {
  float hi, lo;
};
__device__ __forceinline__ Turns to_turns(float x) {
  Turns t;
  t.hi = __fmul_rn(x, kInvTwoPiHi);
  t.lo = __fmaf_rn(x, kInvTwoPiLo, __fmaf_rn(x, kInvTwoPiHi, -t.hi));
  return t;
}
__device__ __forceinline__ void harmonic_sincos(const Turns& t, int k, float& sn, float& cs) {
  const float scale = static_cast<float>(1 << k);
  const float v = t.hi * scale;
  const float r = (v - rintf(v)) + t.lo * scale;  // [-0.5, 0.5] turns
  __sincosf(r * 6.283185307179586f, &sn, &cs);
}

template <int kFmt>
__device__ __forceinline__ uint16_t to_half_bits(float x) {
  if (kFmt == 1) {
    __nv_bfloat16 h = __float2bfloat16_rn(x);
    return *reinterpret_cast<uint16_t*>(&h);
  } else {
    __half h = __float2half_rn(x);
    return *reinterpret_cast<uint16_t*>(&h);
  }
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_shared_u16(uint32_t addr, uint16_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}

// shifts the sign bit of x into the mask word (most significant bit first): after 32 calls bit (31 - j) is the sign of
// the j-th value
__device__ __forceinline__ uint32_t push_sign(uint32_t word, float x) {
  return __funnelshift_l(__float_as_uint(x), word, 1);
}
__device__ __forceinline__ void store_mask_half(uint8_t* mask_row, int half, const uint32_t (&w)[4]) {
  *reinterpret_cast<uint4*>(mask_row + half * 16) = make_uint4(w[0], w[1], w[2], w[3]);
}

// Plain 256-wide layers (no head attached): the whole 128-column half is pulled out of TMEM with one wait, converted
// (ReLU fused into the conversion), and only then `before_store` runs -- for half 0 that is the wait until this layer's
// MMAs have stopped reading activation blocks 0,1 -- so the TMEM latency and the conversions sit in the shadow of
// the MMAs that are still running.
template <int kFmt, bool kRelu, typename BeforeStore>
__device__ __forceinline__ void epilogue_half_plain(uint32_t t_addr, int c_lo, uint32_t act_row, uint32_t swz,
                                                    uint8_t* mask_row, BeforeStore&& before_store) {
  uint32_t v[4][32];
#pragma unroll
  for (int q = 0; q < 4; ++q) tmem_ld32(t_addr + c_lo + q * 32, v[q]);
  tmem_ld_wait();
  uint32_t pk[64];
  uint32_t sign[4] = {0u, 0u, 0u, 0u};
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const float x0 = __uint_as_float(v[q][j]), x1 = __uint_as_float(v[q][j + 1]);
      pk[q * 16 + j / 2] = kRelu ? pack_relu<kFmt>(x0, x1) : Half2Pack<kFmt>::pack(x0, x1);
      if (kRelu && mask_row) sign[q] = push_sign(push_sign(sign[q], x0), x1);
    }
  if (kRelu && mask_row) store_mask_half(mask_row, c_lo >> 7, sign);
  before_store();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int c0 = c_lo + q * 32;
    const uint32_t blk = act_row + (c0 >> 6) * kBlkBytes;
    const uint32_t u0 = ((c0 >> 5) & 1) * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      st_shared_v4(blk + (((u0 + i) ^ swz) << 4), pk[q * 16 + 4 * i], pk[q * 16 + 4 * i + 1], pk[q * 16 + 4 * i + 2],
                   pk[q * 16 + 4 * i + 3]);
  }
}

// colour hidden layer (128 wide): + per-ray direction bias (LinearWithRepeat's ray half, fp32), ReLU, 16-bit ->
// activation blocks 0,1 = the A operand of the colour-head MMA
template <int kFmt>
__device__ __forceinline__ void epilogue_color_hidden(uint32_t t_addr, const float* __restrict__ dirbias_row,
                                                      uint32_t act_row, uint32_t swz, uint8_t* mask_row) {
#pragma unroll 1
  for (int cb = 0; cb < 2; ++cb) {  // one 64-column activation block per iteration
    uint32_t v[2][32];
    tmem_ld32(t_addr + cb * 64, v[0]);
    tmem_ld32(t_addr + cb * 64 + 32, v[1]);
    tmem_ld_wait();
    const uint32_t blk = act_row + cb * kBlkBytes;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t pk[16];
      uint32_t sign = 0u;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(dirbias_row + cb * 64 + h * 32 + j));
        const float x0 = __uint_as_float(v[h][j]) + b.x, x1 = __uint_as_float(v[h][j + 1]) + b.y;
        const float x2 = __uint_as_float(v[h][j + 2]) + b.z, x3 = __uint_as_float(v[h][j + 3]) + b.w;
        pk[j / 2] = pack_relu<kFmt>(x0, x1);
        pk[j / 2 + 1] = pack_relu<kFmt>(x2, x3);
        if (mask_row) sign = push_sign(push_sign(push_sign(push_sign(sign, x0), x1), x2), x3);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        st_shared_v4(blk + (((h * 4 + i) ^ swz) << 4), pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
      if (mask_row) *reinterpret_cast<uint32_t*>(mask_row + (cb * 2 + h) * 4) = sign;
    }
  }
}

template <int kFmt, bool kStash>
__global__ void __launch_bounds__(kFwdThreads, 1) mlp_fwd_kernel(const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_act = smem_base + kSmemAct;
  const uint32_t s_emb = smem_base + kSmemEmb;
  const uint32_t s_ring = smem_base + kSmemRing;
  const uint32_t s_bar = smem_base + kSmemBar;
  // barriers (8 B each): full[4], empty[4], then per tile g: half_full[g][2], blk01_free[g], epi_done[g][2],
  // next_pair[g], dens_full[g]; then the TMEM base address
  const uint32_t bar_full = s_bar, bar_empty = s_bar + 8 * kRing, bar_hfull = s_bar + 16 * kRing,
                 bar_b01 = bar_hfull + 32, bar_epi = bar_b01 + 16, bar_next = bar_epi + 32, bar_dfull = bar_next + 16, s_tmem_ptr = bar_dfull + 16;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const Arch& A = p.arch;
  const int L = A.n_mma_layers();
  const int64_t n_tiles = (p.n_points + kTileM - 1) / kTileM;
  const int64_t n_pairs = (n_tiles + 1) / 2;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kRing; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 2);  // both MMA issuers release a weight block
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(bar_hfull + 16 * g, 1);
      mbar_init(bar_hfull + 16 * g + 8, 1);
      mbar_init(bar_b01 + 8 * g, 1);
      mbar_init(bar_epi + 16 * g, 128);
      mbar_init(bar_epi + 16 * g + 8, 128);
      mbar_init(bar_next + 8 * g, 128);
      mbar_init(bar_dfull + 8 * g, 1);
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
    // ---------------------------------------------------------------- TMA producer
    if (elect_one()) {
      uint32_t slot = 0, phase = 0;
      const bool acct = kInstrument && p.trace && blockIdx.x == 0 && (p.debug & 128);
      long long a_empty = 0, a_t0 = acct ? clock64() : 0;
      for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        int s = 0;
        for (int l = 0; l < L; ++l) {
          if (A.merged(l)) continue;  // the intermediate layer lives inside the colour hidden layer's weights
          const int per_half = A.stages_per_half(l), nkb = A.nkb(l);
          for (int j = 0; j < per_half * A.nnh(l); ++j, ++s) {
            const uint32_t bytes = (j % per_half) == nkb ? kBiasBlkBytes : kBlkBytes;
            YN_TIMED(acct, a_empty, mbar_wait(bar_empty + 8 * slot, phase ^ 1));
            if (debug_flags(p.debug) & 2) {
              mbar_arrive(bar_full + 8 * slot);
            } else {
              mbar_arrive_expect_tx(bar_full + 8 * slot, bytes);
              bulk_g2s(s_ring + slot * kBlkBytes, p.wpack + (size_t)s * kBlkBytes, bytes, bar_full + 8 * slot);
            }
            if (++slot == kRing) { slot = 0; phase ^= 1; }
          }
        }
        // density head stage (8 KB), then colour head stage (4 KB)
        for (int hs = 0; hs < 2; ++hs) {
          const uint32_t bytes = hs == 0 ? kDensBlkBytes : kHeadBlkBytes;
          mbar_wait(bar_empty + 8 * slot, phase ^ 1);
          if (debug_flags(p.debug) & 2) {
            mbar_arrive(bar_full + 8 * slot);
          } else {
            mbar_arrive_expect_tx(bar_full + 8 * slot, bytes);
            bulk_g2s(s_ring + slot * kBlkBytes, p.wpack + (size_t)(A.density_stage() + hs) * kBlkBytes, bytes, bar_full + 8 * slot);
          }
          if (++slot == kRing) { slot = 0; phase ^= 1; }
        }
      }
      if (acct) {
        long long* o = p.trace + (size_t)3 * 2 * kTraceEvents;
        o[0] = 0x7a11;
        o[1] = clock64() - a_t0;
        o[2] = a_empty;
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == 10) {
    // ---------------------------------------------------------------- MMA issuers (one per tile of the pair)
    // Both issuers consume every weight block (one L2 read feeds both tiles); each owns its tile's accumulators
    // and barriers.  Dependencies on the tile's epilogue group:
    //   epi_done[g][0]: output blocks 0,1 of the previous layer written and accumulator half 0 drained
    //   epi_done[g][1]: blocks 2,3 written and accumulator half 1 drained
    const int g = warp == 1 ? 0 : 1;
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc(128, 128, kFmt, 0, 0);
      const uint32_t my_epi = bar_epi + 16 * g, my_hfull = bar_hfull + 16 * g, my_b01 = bar_b01 + 8 * g;
      const uint32_t act_base = s_act + g * 4 * kBlkBytes, emb_base = s_emb + g * kBlkBytes;
      uint32_t slot = 0, phase = 0;
      uint32_t ed_phase0 = 0, ed_phase1 = 0, next_phase = 0;
      const uint32_t my_next = bar_next + 8 * g;
      Tracer tr;
      const bool acct = kInstrument && p.trace && blockIdx.x == 0 && g == 0 && (p.debug & 128);
      tr.init((kInstrument && (p.debug & 128)) ? nullptr : p.trace, g);
      long long a_epi0 = 0, a_epi1 = 0, a_full = 0, a_t0 = acct ? clock64() : 0;
      for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        for (int l = 0; l < L; ++l) {
          if (A.merged(l)) continue;
          const int nkbh = A.nkb_hidden(l);
          const int nkb = A.nkb(l);
          const int nnh = A.nnh(l);
          const int kb_free = nkb > 1 ? 1 : 0;
          const bool has_bias = A.has_bias_stage(l);
          tr.log(l << 8 | 0);
          if (l == 0 && pair != (int64_t)blockIdx.x) {
            // later pairs: "embedding ready, accumulator half 0 drained" comes from the previous pair's colour layer on
            // its own barrier (a second completion of epi_done[0] right behind the head's could alias its phase)
            YN_TIMED(acct, a_epi0, mbar_wait(my_next, next_phase));
            next_phase ^= 1;
          } else {
            YN_TIMED(acct, a_epi0, mbar_wait(my_epi, ed_phase0));
            ed_phase0 ^= 1;
          }
          tc_fence_after();
          tr.log(l << 8 | 1);
          bool waited1 = false;
          for (int nh = 0; nh < nnh; ++nh) {
            const uint32_t d_tmem = tmem_base + g * 256 + nh * 128;
            for (int kb = 0; kb < nkb; ++kb) {
              if (!waited1 && (nh == 1 || (kb >= 2 && kb < nkbh))) {
                tr.log(l << 8 | 2);
                YN_TIMED(acct, a_epi1, mbar_wait(my_epi + 8, ed_phase1));
                ed_phase1 ^= 1;
                tc_fence_after();
                waited1 = true;
                tr.log(l << 8 | 3);
              }
              YN_TIMED(acct, a_full, mbar_wait(bar_full + 8 * slot, phase));
              tc_fence_after();
              tr.log(l << 8 | 16 | (nh << 3) | kb);
              const uint64_t b_desc = umma_desc_kmajor(s_ring + slot * kBlkBytes);
              const uint64_t a_desc = umma_desc_kmajor((kb < nkbh) ? (act_base + kb * kBlkBytes) : emb_base);
#pragma unroll
              for (int k = 0; k < 4; ++k)  // K advances by 32 bytes = 2 descriptor units inside the swizzle atom
                umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
              umma_commit(bar_empty + 8 * slot);
              if (++slot == kRing) { slot = 0; phase ^= 1; }
              // activation blocks 0,1 are not read again in this layer after (last half, kb_free); layer 0 reads
              // only the embedding block, so they are free from its first MMA on
              // (the colour hidden layer is followed by the density head, which reads all four blocks again)
              if (l != L - 1 && (l == 0 ? (nh == 0 && kb == 0) : (nh == nnh - 1 && kb == kb_free))) umma_commit(my_b01);
            }
            if (has_bias) {
              // + bias: embedding slice 3 (channel 63 == 1) x the [128 x 16] bias block
              mbar_wait(bar_full + 8 * slot, phase);
              tc_fence_after();
              umma_f16(d_tmem, umma_desc_kmajor(emb_base + 96), umma_desc_kmajor_k16_nosw(s_ring + slot * kBlkBytes), idesc, 1);
              umma_commit(bar_empty + 8 * slot);
              if (++slot == kRing) { slot = 0; phase ^= 1; }
            }
            umma_commit(my_hfull + 8 * nh);
            tr.log(l << 8 | 4 | nh);
          }
          if (l == L - 1) {
            // density head: the same activations (last trunk output, blocks 0-3) x w_d -> 16 accumulator columns
            // [144, 160) of the tile (the colour hidden layer is 128 wide and leaves the second half free)
            constexpr uint32_t idesc_dens = umma_idesc(128, kHeadN, kFmt, 0, 0);
            if (!waited1) {
              mbar_wait(my_epi + 8, ed_phase1);
              ed_phase1 ^= 1;
              tc_fence_after();
              waited1 = true;
            }
            mbar_wait(bar_full + 8 * slot, phase);
            tc_fence_after();
            const uint32_t d_dens = tmem_base + g * 256 + 144;
#pragma unroll
            for (int kb = 0; kb < 4; ++kb) {
              const uint64_t a_desc = umma_desc_kmajor(act_base + kb * kBlkBytes);
              const uint64_t b_desc = umma_desc_kmajor(s_ring + slot * kBlkBytes + kb * 2048);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16(d_dens, a_desc + 2 * k, b_desc + 2 * k, idesc_dens, (kb | k) != 0);
              if (kb == 1) umma_commit(my_b01);  // blocks 0,1 are not read again
            }
            umma_commit(bar_empty + 8 * slot);
            if (++slot == kRing) { slot = 0; phase ^= 1; }
            umma_commit(bar_dfull + 8 * g);
          }
          if (!waited1) {
            mbar_wait(my_epi + 8, ed_phase1);
            ed_phase1 ^= 1;
          }
        }
        // colour head: [128 x 128] hidden activations (blocks 0,1, written by the colour layer's epilogue) x W2^T
        // -> 16 accumulator columns in the tile's (drained) second half; signalled on the "half 1" barrier
        {
          constexpr uint32_t idesc_head = umma_idesc(128, kHeadN, kFmt, 0, 0);
          tr.log(L << 8 | 0);
          mbar_wait(my_epi, ed_phase0);
          ed_phase0 ^= 1;
          mbar_wait(my_epi + 8, ed_phase1);
          ed_phase1 ^= 1;
          tr.log(L << 8 | 1);
          mbar_wait(bar_full + 8 * slot, phase);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + g * 256 + 128;
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t a_desc = umma_desc_kmajor(act_base + kb * kBlkBytes);
            const uint64_t b_desc = umma_desc_kmajor(s_ring + slot * kBlkBytes + kb * 2048);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc_head, (kb | k) != 0);
          }
          umma_commit(bar_empty + 8 * slot);
          if (++slot == kRing) { slot = 0; phase ^= 1; }
          umma_commit(my_hfull + 8);
          tr.log(L << 8 | 5);
        }
      }
      if (acct) {
        long long* o = p.trace;
        o[0] = 0x7a11;
        o[1] = clock64() - a_t0;
        o[2] = a_epi0;
        o[3] = a_epi1;
        o[4] = a_full;
      }
    }
    __syncwarp();
  } else if (warp >= 2 && warp <= 9) {
    // ---------------------------------------------------------------- epilogue groups
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const uint32_t my_epi = bar_epi + 16 * g, my_hfull = bar_hfull + 16 * g, my_b01 = bar_b01 + 8 * g;
    const int row = q * 32 + lane;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * 256;
    const uint32_t act_g = s_act + g * 4 * kBlkBytes;
    const uint32_t emb_g = s_emb + g * kBlkBytes;
    const uint32_t swz = static_cast<uint32_t>(row & 7);
    const uint32_t act_row = act_g + static_cast<uint32_t>(row) * 128u;
    const bool stash_leader = (warp - 2) % 4 == 0 && lane == 0;
    const int nfx = A.n_freq_xyz;
    const int blocks_per_tile = A.stash_blocks_per_tile();
    uint32_t hf_phase0 = 0, hf_phase1 = 0, b01_phase = 0, df_phase = 0;
    Tracer tr;
    const bool acct = kInstrument && p.trace && blockIdx.x == 0 && g == 0 && q == 0 && lane == 0 && (p.debug & 128);
    tr.init(q == 0 && lane == 0 && !(kInstrument && (p.debug & 128)) ? p.trace : nullptr, 2 + g);
    long long a_hf0 = 0, a_b01 = 0, a_hf1 = 0, a_head = 0, a_t0 = acct ? clock64() : 0;
    int l_emb_last = 0;  // last layer that reads the embedding block as an operand
    for (int l = 1; l < A.n_layers; ++l)
      if (A.has_emb(l)) l_emb_last = l;

    // harmonic embedding (models/utils.py:90-103) of row `row` of tile `t` -> this tile's embedding block:
    // [sin(x f_k) | cos(x f_k) | x], channel a*L+k, channel 63 = 1 (bias).  x * 2^k is exact in fp32, so the phase
    // is reduced once per coordinate (see harmonic_sincos); no error growth with the octave.
    // (A double-angle recurrence from the base octave was tried: its error triples per octave, 8e-4 at 2^9 -- rejected.)
    auto write_embedding = [&](int64_t t) {
      const int64_t gi = t * kTileM + row;
      const bool ok = gi < p.n_points;
      float pt[3] = {0.f, 0.f, 0.f};
      if (ok) {
        const int64_t r = gi / p.P;
        const float z = __ldg(p.lengths + gi);
#pragma unroll
        for (int a = 0; a < 3; ++a)
          pt[a] = __fadd_rn(__ldg(p.origins + r * 3 + a), __fmul_rn(z, __ldg(p.directions + r * 3 + a)));
      }
      if (nfx == 10) {
        // lego / fern shape: whole 64-channel row assembled in registers, eight conflict-free 16-byte stores
        uint32_t pk[32];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const Turns turns = to_turns(pt[a]);
#pragma unroll
          for (int k = 0; k < 10; k += 2) {
            float s0, c0, s1, c1;
            harmonic_sincos(turns, k, s0, c0);
            harmonic_sincos(turns, k + 1, s1, c1);
            if (!ok) { s0 = s1 = c0 = c1 = 0.f; }
            pk[(a * 10 + k) >> 1] = Half2Pack<kFmt>::pack(s0, s1);
            pk[(30 + a * 10 + k) >> 1] = Half2Pack<kFmt>::pack(c0, c1);
          }
        }
        pk[30] = Half2Pack<kFmt>::pack(pt[0], pt[1]);
        pk[31] = Half2Pack<kFmt>::pack(pt[2], 1.f);
        const uint32_t row_base = emb_g + static_cast<uint32_t>(row) * 128u;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          st_shared_v4(row_base + ((static_cast<uint32_t>(c) ^ swz) << 4), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        return;
      }
      for (int a = 0; a < 3; ++a) {
        const Turns turns = to_turns(pt[a]);
        for (int k = 0; k < nfx; ++k) {
          float sn, cs;
          harmonic_sincos(turns, k, sn, cs);
          if (!ok) { sn = 0.f; cs = 0.f; }
          const int ch = a * nfx + k;
          st_shared_u16(emb_g + sw128_offset(row, ch), to_half_bits<kFmt>(sn));
          st_shared_u16(emb_g + sw128_offset(row, 3 * nfx + ch), to_half_bits<kFmt>(cs));
        }
        st_shared_u16(emb_g + sw128_offset(row, 6 * nfx + a), to_half_bits<kFmt>(pt[a]));
      }
      for (int ch = 6 * nfx + 3; ch < 63; ++ch) st_shared_u16(emb_g + sw128_offset(row, ch), 0);
      st_shared_u16(emb_g + sw128_offset(row, 63), to_half_bits<kFmt>(1.f));
    };

    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      const int64_t tile = 2 * pair + g;
      const int64_t gidx = tile * kTileM + row;
      const bool valid = gidx < p.n_points;
      const int64_t ray = valid ? gidx / p.P : 0;
      // (YN_FWD_DEBUG bit 4, timing experiment: every stash store lands in a 64-tile window that stays in L2)
      uint8_t* stash_tile = kStash ? p.stash + (size_t)((debug_flags(p.debug) & 4) ? tile % 64 : tile) * blocks_per_tile * kBlkBytes : nullptr;
      const bool tile_live = tile < n_tiles && !(kStash && (debug_flags(p.debug) & 16));  // (bit 16, timing experiment: no stash stores)

      if (kStash) {
        if (stash_leader) bulk_wait_read<0>();
        named_bar_sync(1 + g, 128);
      }
      // ---- embedding of this tile: computed here for the CTA's first pair, otherwise prefetched during the
      // previous pair (see below)
      const bool first_pair = pair == (int64_t)blockIdx.x;
      if (first_pair) write_embedding(tile);
      // the embedding acts as the epilogue of a virtual layer -1.  "Half 0 done" (embedding written, accumulator
      // columns [0,128) drained) was already signalled from the previous pair's colour layer, so that layer 0 starts
      // under the colour head; "half 1 done" (the head's 16 columns have been read) is signalled here.
      tc_fence_before();
      fence_proxy_async_smem();
      if (first_pair) mbar_arrive(my_epi);
      mbar_arrive(my_epi + 8);
      if (kStash) {
        named_bar_sync(1 + g, 128);
        if (stash_leader && tile_live) {
          bulk_s2g(stash_tile, emb_g, kBlkBytes);
          bulk_commit();
        }
      }

      for (int l = 0; l < L; ++l) {
        if (A.merged(l)) continue;
        const bool is_color = (l == L - 1);
        const bool is_last_trunk = (l == L - 3);
        const float* bias = is_color ? p.dirbias + ray * kDirPad : p.aux + A.aux_bias(l);
        if (is_last_trunk) {  // the colour layer's per-ray bias row (512 B): pull it into L1 one layer ahead
          const float* row_bias = p.dirbias + ray * kDirPad;
#pragma unroll
          for (int i = 0; i < 4; ++i) asm volatile("prefetch.global.L1 [%0];" ::"l"(row_bias + 32 * i));
        }
        // ---- half 0: accumulator columns [0,128) -> activation blocks 0,1
        tr.log(l << 8 | 0);
        YN_TIMED(acct, a_hf0, mbar_wait(my_hfull, hf_phase0));
        hf_phase0 ^= 1;
        tr.log(l << 8 | 1);
        tc_fence_after();
        auto before_store0 = [&] {
          YN_TIMED(acct, a_b01, mbar_wait(my_b01, b01_phase));  // this layer's MMAs no longer read blocks 0,1
          b01_phase ^= 1;
          tc_fence_after();
          tr.log(l << 8 | 2);
          if (kStash) {
            // the previous layer's stash store of blocks 0,1 has read the buffer (its store of blocks 2,3, the most
            // recent group, may still be in flight)
            if (stash_leader) bulk_wait_read<1>();
            named_bar_sync(1 + g, 128);
          }
        };
        auto before_store1 = [&] {
          if (kStash) {  // ... and the one of blocks 2,3 (most recent group now: this layer's blocks 0,1)
            if (stash_leader) bulk_wait_read<1>();
            named_bar_sync(1 + g, 128);
          }
        };
        const bool plain = !is_color;
        // ReLU sign mask of this layer and row (training only): trunk layer l -> mask l, colour hidden -> mask n_layers
        uint8_t* mask_row = nullptr;
        if (kStash && tile_live && !(debug_flags(p.debug) & 8))  // (bit 8, timing experiment: no sign masks)
          mask_row = stash_tile + A.mask_offset(is_color ? A.n_layers : l) + (size_t)row * 32;
        if (debug_flags(p.debug) & 1) {
          before_store0();
        } else if (plain) {
          epilogue_half_plain<kFmt, true>(t_row, 0, act_row, swz, mask_row, before_store0);
        } else if (is_color) {
          before_store0();
          epilogue_color_hidden<kFmt>(t_row, bias, act_row, swz, mask_row);
        }
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(my_epi);
        tr.log(l << 8 | 3);
        if (kStash) {
          // stash blocks 0,1 right away (half a layer earlier than blocks 2,3: the stores are spread over time instead of
          // arriving as one 64 KB burst per layer)
          named_bar_sync(1 + g, 128);
          if (stash_leader && tile_live) {
            uint8_t* dst = stash_tile + (size_t)A.stash_block_of_layer(l) * kBlkBytes;
            bulk_s2g(dst, act_g, kBlkBytes);
            bulk_s2g(dst + kBlkBytes, act_g + kBlkBytes, kBlkBytes);
          }
          if (stash_leader) bulk_commit();
        }
        // ---- half 1: columns [128,256) -> blocks 2,3 (the colour hidden layer is 128 wide: nothing to do)
        if (!is_color) {
          YN_TIMED(acct, a_hf1, mbar_wait(my_hfull + 8, hf_phase1));
          hf_phase1 ^= 1;
          tc_fence_after();
          tr.log(l << 8 | 4);
          if (debug_flags(p.debug) & 1) {
          } else
            epilogue_half_plain<kFmt, true>(t_row, 128, act_row, swz, mask_row, before_store1);
          tc_fence_before();
          fence_proxy_async_smem();
        }
        mbar_arrive(my_epi + 8);  // (colour layer: blocks 0,1 hold the hidden activations, the head MMA may start)
        tr.log(l << 8 | 5);
        if (is_color && pair + gridDim.x < n_pairs) {
          // next pair, layer 0, half 0 may start right behind the head MMA: its embedding was prefetched and
          // accumulator columns [0,128) are drained
          mbar_arrive(bar_next + 8 * g);
        }
        if (l == l_emb_last && pair + gridDim.x < n_pairs) {
          // every MMA that reads this tile's embedding as an operand has completed (half_full[1] of this layer):
          // prefetch the next pair's embedding now, in the shadow of the remaining layers.  Later bias MMAs read
          // only channels 48..63 of it against zero weights (and the constant 1 of channel 63 is rewritten as 1).
          if (kStash) {
            if (stash_leader) bulk_wait_read<0>();
            named_bar_sync(1 + g, 128);
          }
          write_embedding(2 * (pair + gridDim.x) + g);
        }
        if (is_color) {
          // density head: accumulator column 144 of the tile
          YN_TIMED(acct, a_head, mbar_wait(bar_dfull + 8 * g, df_phase));
          df_phase ^= 1;
          tc_fence_after();
          uint32_t dv[4];
          tmem_ld4(t_row + 144, dv);
          tmem_ld_wait();
          if (valid) p.density[gidx] = __uint_as_float(dv[0]) + __ldg(p.aux + A.aux_bd());
          // colour head: 16 accumulator columns at the start of the tile's second half (the first color_dim are real);
          // the next pair's embedding arrival (program order) covers the TMEM hand-over
          YN_TIMED(acct, a_head, mbar_wait(my_hfull + 8, hf_phase1));
          hf_phase1 ^= 1;
          tc_fence_after();
          uint32_t hv[4];
          tmem_ld4(t_row + 128, hv);
          tmem_ld_wait();
          tr.log(L << 8 | 4);
          if (valid) {
            const int C = A.color_dim;
#pragma unroll
            for (int c = 0; c < 3; ++c)
              if (c < C) {
                const float x = __uint_as_float(hv[c]) + __ldg(p.aux + A.aux_b2() + c);
                p.rgb[gidx * C + c] = 1.f / (1.f + expf(-x));
              }
          }
        }
        if (kStash && !is_color) {
          named_bar_sync(1 + g, 128);
          if (stash_leader && tile_live) {
            uint8_t* dst = stash_tile + (size_t)(A.stash_block_of_layer(l) + 2) * kBlkBytes;
            bulk_s2g(dst, act_g + 2 * kBlkBytes, kBlkBytes);
            bulk_s2g(dst + kBlkBytes, act_g + 3 * kBlkBytes, kBlkBytes);
          }
          if (stash_leader) bulk_commit();
        }
      }
    }
    if (kStash && stash_leader) bulk_wait<0>();
    if (acct) {
      long long* o = p.trace + (size_t)2 * 2 * kTraceEvents;
      o[0] = 0x7a11;
      o[1] = clock64() - a_t0;
      o[2] = a_hf0;
      o[3] = a_b01;
      o[4] = a_hf1;
      o[5] = a_head;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static int launch_fwd(const FwdParams& p, cudaStream_t stream) {
  const int64_t n_tiles = (p.n_points + kTileM - 1) / kTileM;
  const int64_t n_pairs = (n_tiles + 1) / 2;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)(n_pairs < sms ? n_pairs : sms);
  if (grid == 0) return YN_OK;
  auto run = [&](auto kern) {
    // once per kernel (all instantiations share one pointer type, so key on the pointer); not a stream operation,
    // and kept out of stream capture this way
    if (first_use(reinterpret_cast<const void*>(kern)))
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes);
    kern<<<grid, kFwdThreads, kFwdSmemBytes, stream>>>(p);
  };
  const bool st = p.stash != nullptr;
  if (p.arch.fmt == 1) {
    if (st) run(mlp_fwd_kernel<1, true>); else run(mlp_fwd_kernel<1, false>);
  } else {
    if (st) run(mlp_fwd_kernel<0, true>); else run(mlp_fwd_kernel<0, false>);
  }
  return check_launch("yn_mlp_fwd");
}

}  // namespace ynb

extern "C" int yn_mlp_fwd(const yn_mlp_arch* arch, const float* origins, const float* directions,
                          const float* lengths, const float* dirbias, const void* wpack, const float* aux,
                          float* density, float* rgb, void* stash, int64_t R, int P, void* stream) {
  if (int rc = ynb::check_arch(arch)) return rc;
  if (R < 0 || P <= 0) return ynb::fail(YN_ERR_INVALID_ARGUMENT, "yn_mlp_fwd: bad sizes R=%lld P=%d", (long long)R, P);
  if (R == 0) return YN_OK;
  if (!origins || !directions || !lengths || !dirbias || !wpack || !aux || !density || !rgb)
    return ynb::fail(YN_ERR_INVALID_ARGUMENT, "yn_mlp_fwd: null pointer");
  ynb::FwdParams p;
  p.arch = ynb::arch_from_c(arch);
  p.origins = origins;
  p.directions = directions;
  p.lengths = lengths;
  p.dirbias = dirbias;
  p.wpack = static_cast<const uint8_t*>(wpack);
  p.aux = aux;
  p.density = density;
  p.rgb = rgb;
  p.stash = static_cast<uint8_t*>(stash);
  p.n_points = R * P;
  p.P = P;
  {
    const char* dbg = getenv("YN_FWD_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
    static bool warned = false;
    if (!ynb::kInstrument && !warned && (p.debug || getenv("YN_FWD_TRACE"))) {
      warned = true;
      fprintf(stderr, "yn_mlp_fwd: YN_FWD_DEBUG / YN_FWD_TRACE need a `make INSTRUMENT=1` build of the library; ignored\n");
    }
  }
  p.trace = nullptr;
  if (const char* path = getenv("YN_FWD_TRACE")) {  // debug aid: synchronous, one launch per file
    const size_t n = (size_t)4 * 2 * ynb::kTraceEvents;
    long long* dev = nullptr;
    cudaMalloc(&dev, n * sizeof(long long));
    cudaMemset(dev, 0, n * sizeof(long long));
    p.trace = dev;
    const int rc = ynb::launch_fwd(p, static_cast<cudaStream_t>(stream));
    cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
    long long* host = (long long*)malloc(n * sizeof(long long));
    cudaMemcpy(host, dev, n * sizeof(long long), cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(path, "wb")) {
      fwrite(host, sizeof(long long), n, f);
      fclose(f);
    }
    free(host);
    cudaFree(dev);
    return rc;
  }
  return ynb::launch_fwd(p, static_cast<cudaStream_t>(stream));
}
