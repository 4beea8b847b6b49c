// RayPointRefiner + sample_pdf in one launch, one warp per ray.
//
// Replaces RayPointRefiner.forward and sample_pdf_python
// (yanerf/pipelines/renderers/utils.py:48-69, 83-158): midpoints of the depth samples, PDF -> CDF,
// searchsorted(right=True), inverse-CDF interpolation, concatenation with the input depths and an ascending
// sort.
//
// Bit-exactness: the reference's searchsorted indices depend on the rounding of `sum` and `cumsum`.  The
// parity oracle is the reference's --device cpu path, so this kernel reproduces torch-CPU's fp32 orders:
//   * Tensor.sum(-1) over a contiguous row = ATen's vectorised inner reduction: 8 vector lanes, 4 interleaved
//     accumulators per lane with a 4-level cascade every 16 rows, leftover vectors into accumulator 0,
//     accumulators 1..3 folded into 0, then the scalar tail summed from zero, then the 8 lanes in order;
//   * cumsum = fp64 running sum, each prefix rounded to fp32;
//   * every other op is a single correctly rounded fp32 operation (no FMA contraction).
// The fp64 prefix is computed with a parallel scan whose result can differ from the sequential fp64 sum by at most
// (K-1) 2^-53 relative (< 1e-14 for K < 100); whenever a prefix lands within 4e-14 relative of an fp32 rounding
// boundary (where that difference could change the rounded value) the warp redoes that ray sequentially.
#include <cuda_runtime.h>
#include <math_constants.h>

#include "mlp_common.cuh"
#include "rng.cuh"

namespace ynb {

struct PdfParams {
  const float* z;
  const float* w;
  const float* u;    // nullptr: U[0,1) draws generated in the kernel from `rng`
  RngRef rng;
  int64_t u_stride;  // 0: one shared row of draws (deterministic linspace), n_new: per-ray draws
  float* out;
  int64_t* inds;
  int32_t* flag;
  int64_t R;
  int P;
  int n_new;
  int add_input;
  int sort_pow2;  // power of two >= P + n_new
  int bins_mode;  // 1: `z` already holds the bin edges [R,P] and `w` the P-1 bin weights (plain sample_pdf)
  int sort_out;   // 0: keep the samples in draw order (bins_mode only)
};

// partial sum of vector lane j (0..VL-1) in ATen's cascade order over wp[0..K); VL = 8 is the vectorised inner
// reduction, VL = 1 the scalar row_sum ATen falls back to when the row is shorter than one vector
template <int VL>
__device__ float aten_lane_partial(const float* wp, int K, int j) {
  const int vec_size = K / VL;
  const int size_ilp = vec_size / 4;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  int ceil_log2 = 0;
  while ((1 << ceil_log2) < size_ilp) ++ceil_log2;
  int level_power = ceil_log2 / 4;
  if (level_power < 4) level_power = 4;
  const int level_step = 1 << level_power;
  const int level_mask = level_step - 1;
  int i = 0;
  for (; i + level_step <= size_ilp;) {
    for (int jj = 0; jj < level_step; ++jj, ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[0][c] = __fadd_rn(acc[0][c], wp[(i * 4 + c) * VL + j]);
#pragma unroll
    for (int lev = 1; lev < 4; ++lev) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        acc[lev][c] = __fadd_rn(acc[lev][c], acc[lev - 1][c]);
        acc[lev - 1][c] = 0.f;
      }
      const int mask = level_mask << (lev * level_power);
      if ((i & mask) != 0) break;
    }
  }
  for (; i < size_ilp; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[0][c] = __fadd_rn(acc[0][c], wp[(i * 4 + c) * VL + j]);
#pragma unroll
  for (int lev = 1; lev < 4; ++lev)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[0][c] = __fadd_rn(acc[0][c], acc[lev][c]);
  float part = acc[0][0];
  for (int v = size_ilp * 4; v < vec_size; ++v) part = __fadd_rn(part, wp[v * VL + j]);
  part = __fadd_rn(part, acc[0][1]);
  part = __fadd_rn(part, acc[0][2]);
  part = __fadd_rn(part, acc[0][3]);
  return part;
}

__device__ __forceinline__ double warp_scan_incl_d(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// bitonic sort of n (power of two) floats in shared memory by one warp
__device__ void warp_bitonic_sort(float* a, int n, int lane) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (n >> 1); t += 32) {
        const int lo = ((t / j) * 2 * j) + (t % j);
        const int hi = lo + j;
        const bool up = ((lo & k) == 0);
        const float x = a[lo], y = a[hi];
        if ((x > y) == up) {
          a[lo] = y;
          a[hi] = x;
        }
      }
      __syncwarp();
    }
  }
}

// bitonic sort of 32*NPL floats held in registers, element e = lane*NPL + r; partners inside a lane are
// compare-swapped in registers, partners in other lanes through one shuffle per element
template <int NPL>
__device__ __forceinline__ void warp_bitonic_sort_regs(float (&v)[NPL], int lane) {
  constexpr int N = 32 * NPL;
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j < NPL) {
#pragma unroll
        for (int r = 0; r < NPL; ++r) {
          const int pr = r ^ j;
          if (pr > r) {
            const bool up = ((lane * NPL + r) & k) == 0;
            const float a = v[r], b = v[pr];
            const bool sw = (a > b) == up;
            v[r] = sw ? b : a;
            v[pr] = sw ? a : b;
          }
        }
      } else {
        const int lj = j / NPL;
        const bool lower = (lane & lj) == 0;
#pragma unroll
        for (int r = 0; r < NPL; ++r) {
          const float other = __shfl_xor_sync(0xffffffffu, v[r], lj);
          const bool up = ((lane * NPL + r) & k) == 0;
          v[r] = (lower == up) ? fminf(v[r], other) : fmaxf(v[r], other);
        }
      }
    }
  }
}

// sorts s_new[0 .. 32*NPL) ascending (skipped when already sorted, the usual case for deterministic draws)
template <int NPL>
__device__ __forceinline__ void sort_new_samples(float* s_new, int lane) {
  float v[NPL];
#pragma unroll
  for (int r = 0; r < NPL; ++r) v[r] = s_new[lane * NPL + r];
  bool ok = true;
#pragma unroll
  for (int r = 0; r + 1 < NPL; ++r) ok &= v[r] <= v[r + 1];
  const float nxt = __shfl_down_sync(0xffffffffu, v[0], 1);
  ok &= (lane == 31) || (v[NPL - 1] <= nxt);
  if (__all_sync(0xffffffffu, ok)) return;
  warp_bitonic_sort_regs<NPL>(v, lane);
#pragma unroll
  for (int r = 0; r < NPL; ++r) s_new[lane * NPL + r] = v[r];
  __syncwarp();
}

__global__ void __launch_bounds__(256) sample_pdf_merge_kernel(const PdfParams p) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  if (ray >= p.R) return;
  const int P = p.P, N = p.n_new;
  const int NB = p.bins_mode ? P : P - 1;  // bin edges
  const int K = NB - 1;                    // pdf weights
  const int per_warp = 3 * P + p.sort_pow2;
  float* s_cdf = smem + (size_t)wib * per_warp;  // [NB] : first holds w + eps, then the cdf
  float* s_bin = s_cdf + P;                      // [NB]
  float* s_z = s_bin + P;                        // [P]
  float* s_new = s_z + P;                        // [sort_pow2]
  const float* zr = p.z + ray * P;
  // refiner mode: weights[:, 1:-1] of the [R,P] raymarcher weights; bins mode: the [R,P-1] row as is
  const float* wr = p.bins_mode ? p.w + ray * (int64_t)(P - 1) - 1 : p.w + ray * P;

  // ---- load: z, midpoints (torch.lerp(z[1:], z[:-1], 0.5) = b - (b - a) * 0.5), weights + eps
  bool bad = false, unsorted = false;
  for (int i = lane; i < P; i += 32) {
    const float zi = __ldg(zr + i);
    s_z[i] = zi;
    if (p.bins_mode) {
      s_bin[i] = zi;
    } else if (i + 1 < P) {
      const float zn = __ldg(zr + i + 1);
      s_bin[i] = __fsub_rn(zi, __fmul_rn(__fsub_rn(zi, zn), 0.5f));
      unsorted |= zn < zi;
    }
    if (i >= 1 && i <= K) {
      const float wv = __fadd_rn(__ldg(wr + i), 1e-5f);
      s_cdf[i] = wv;  // wp[k] lives at s_cdf[k + 1]
      bad |= wv <= 0.f;  // like `weights.min() <= 0` (a NaN weight does not raise in the reference either)
    }
  }
  unsorted = __any_sync(0xffffffffu, unsorted);
  bad = __any_sync(0xffffffffu, bad);
  if (bad && lane == 0) atomicOr(p.flag, 1);
  __syncwarp();
  const float* wp = s_cdf + 1;

  // ---- sum in ATen order
  float total = 0.f;
  if (K >= 8) {
    float part = 0.f;
    if (lane < 8) part = aten_lane_partial<8>(wp, K, lane);
    for (int k = (K / 8) * 8; k < K; ++k) total = __fadd_rn(total, wp[k]);
#pragma unroll
    for (int j = 0; j < 8; ++j) total = __fadd_rn(total, __shfl_sync(0xffffffffu, part, j));
  } else {
    if (lane == 0) total = aten_lane_partial<1>(wp, K, 0);
    total = __shfl_sync(0xffffffffu, total, 0);
  }

  // ---- pdf and cdf (fp64 running sum rounded per prefix)
  double carry = 0.0;
  bool ambiguous = false;
  const int nch = (K + 31) / 32;
  for (int ch = 0; ch < nch; ++ch) {
    const int k = ch * 32 + lane;
    const float pdf = k < K ? __fdiv_rn(wp[k], total) : 0.f;
    const double incl = warp_scan_incl_d(static_cast<double>(pdf), lane) + carry;
    carry = __shfl_sync(0xffffffffu, incl, 31);
    const float f = static_cast<float>(incl);
    if (k < K) {
      const double fu = static_cast<double>(nextafterf(f, CUDART_INF_F));
      const double fd = static_cast<double>(nextafterf(f, -CUDART_INF_F));
      const double mid_up = 0.5 * (static_cast<double>(f) + fu), mid_dn = 0.5 * (static_cast<double>(f) + fd);
      ambiguous |= (mid_up - incl < 4e-14 * incl) || (incl - mid_dn < 4e-14 * incl);
    }
    __syncwarp();
    if (k < K) s_cdf[k + 1] = f;  // overwrites wp[k], already consumed by every lane of this chunk
    // later chunks still read wp[k'] for k' >= (ch + 1) * 32 only
  }
  if (lane == 0) s_cdf[0] = 0.f;
  __syncwarp();
  if (__any_sync(0xffffffffu, ambiguous)) {
    // exact sequential redo (practically never taken)
    if (lane == 0) {
      double run = 0.0;
      for (int k = 0; k < K; ++k) {
        const float wv = __fadd_rn(__ldg(wr + k + 1), 1e-5f);
        run += static_cast<double>(__fdiv_rn(wv, total));
        s_cdf[k + 1] = static_cast<float>(run);
      }
    }
    __syncwarp();
  }

  // ---- inverse-CDF samples
  const float* ur = p.u ? p.u + ray * p.u_stride : nullptr;
  for (int j = lane; j < N; j += 32) {
    const float u = ur ? __ldg(ur + j) : UniformRow(p.rng, ray).get(j);
    int lo = 0, hi = NB;  // first index with cdf[idx] > u
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_cdf[mid] <= u) lo = mid + 1; else hi = mid;
    }
    const int ind = lo;
    const int below = ind - 1 > 0 ? ind - 1 : 0;
    const int above = ind < NB - 1 ? ind : NB - 1;
    const float cb = s_cdf[below], ca = s_cdf[above];
    const float bb = s_bin[below], ba = s_bin[above];
    float den = __fsub_rn(ca, cb);
    if (den < 1e-5f) den = 1.f;
    const float t = __fdiv_rn(__fsub_rn(u, cb), den);
    s_new[j] = __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));
    if (p.inds) p.inds[ray * N + j] = ind;
  }
  __syncwarp();

  // ---- sort / merge
  float* outr = p.out + ray * (int64_t)(N + (p.add_input ? P : 0));
  if (!p.sort_out) {
    for (int j = lane; j < N; j += 32) outr[j] = s_new[j];
    return;
  }
  if (!p.add_input || unsorted) {
    int total_n = N;
    if (p.add_input) {
      for (int i = lane; i < P; i += 32) s_new[N + i] = s_z[i];
      total_n = N + P;
    }
    int n2 = 1;
    while (n2 < total_n) n2 <<= 1;
    for (int i = total_n + lane; i < n2; i += 32) s_new[i] = CUDART_INF_F;
    __syncwarp();
    warp_bitonic_sort(s_new, n2, lane);
    for (int i = lane; i < total_n; i += 32) outr[i] = s_new[i];
    return;
  }
  if (N == 128) {
    sort_new_samples<4>(s_new, lane);
  } else if (N == 64) {
    sort_new_samples<2>(s_new, lane);
  } else {
    int n2 = 1;
    while (n2 < N) n2 <<= 1;
    for (int i = N + lane; i < n2; i += 32) s_new[i] = CUDART_INF_F;
    __syncwarp();
    warp_bitonic_sort(s_new, n2, lane);
  }
  // merge path: rank of every element in the union (coarse first on ties)
  for (int i = lane; i < P; i += 32) {
    const float v = s_z[i];
    int lo = 0, hi = N;  // number of new samples strictly below v
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_new[mid] < v) lo = mid + 1; else hi = mid;
    }
    outr[i + lo] = v;
  }
  for (int j = lane; j < N; j += 32) {
    const float v = s_new[j];
    int lo = 0, hi = P;  // number of coarse samples <= v
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_z[mid] <= v) lo = mid + 1; else hi = mid;
    }
    outr[j + lo] = v;
  }
}

}  // namespace ynb

#include "sample_pdf_fast.cuh"

template <int PS, int NPL, bool kRng>
static void launch_fast_impl(const ynb::PdfParams& p, cudaStream_t st) {
  constexpr int P = 32 * PS;
  constexpr int CDFN = P <= 64 ? 64 : (P <= 128 ? 128 : 256);
  const int wpb = 8;
  const size_t smem = (size_t)wpb * (CDFN + 2 * P + p.sort_pow2) * sizeof(float);
  auto kern = ynb::sample_pdf_merge_fast_kernel<PS, NPL, kRng>;
  static size_t configured = 0;  // per instantiation; the size only depends on (PS, NPL)
  if (configured < smem) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  kern<<<(unsigned)((p.R + wpb - 1) / wpb), wpb * 32, smem, st>>>(p);
}

template <int PS, int NPL>
static void launch_fast(const ynb::PdfParams& p, cudaStream_t st) {
  if (p.u == nullptr) launch_fast_impl<PS, NPL, true>(p, st);  // draws generated in the kernel
  else launch_fast_impl<PS, NPL, false>(p, st);
}

static int launch_pdf(const float* lengths, const float* weights, const float* u, int64_t u_row_stride,
                      const int64_t* rng_state, int rng_site, float* new_lengths, int64_t* inds, int32_t* flag, int64_t R, int P, int n_new,
                      int add_input_samples, int bins_mode, int sort_out, void* stream) {
  if (R < 0 || P < 3 || n_new < 1)
    return ynb::fail(YN_ERR_INVALID_ARGUMENT, "yn_sample_pdf_merge: bad sizes R=%lld P=%d n_new=%d", (long long)R, P, n_new);
  if (R == 0) return YN_OK;
  if (!lengths || !weights || (!u && !rng_state) || !new_lengths || !flag)
    return ynb::fail(YN_ERR_INVALID_ARGUMENT, "yn_sample_pdf_merge: null pointer");
  if (u && u_row_stride != 0 && u_row_stride < n_new)
    return ynb::fail(YN_ERR_INVALID_ARGUMENT, "yn_sample_pdf_merge: u_row_stride must be 0 or >= n_new");
  ynb::PdfParams p;
  p.z = lengths; p.w = weights; p.u = u; p.u_stride = u ? u_row_stride : 0; p.out = new_lengths; p.inds = inds; p.flag = flag;
  p.rng.state = rng_state; p.rng.site = rng_site;
  p.R = R; p.P = P; p.n_new = n_new; p.add_input = add_input_samples;
  int n2 = 1;
  while (n2 < P + n_new) n2 <<= 1;
  p.sort_pow2 = n2;
  p.bins_mode = bins_mode;
  p.sort_out = sort_out;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const auto al8 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 7u) == 0; };
  if (!bins_mode && add_input_samples && sort_out && al8(lengths) && al8(weights) && al8(u) && al8(new_lengths) &&
      (p.u_stride % 2 == 0)) {
    const int key = P * 1000 + n_new;
    bool done = true;
    switch (key) {
      case 64128: launch_fast<2, 4>(p, st); break;
      case 64064: launch_fast<2, 2>(p, st); break;
      case 128128: launch_fast<4, 4>(p, st); break;
      case 128064: launch_fast<4, 2>(p, st); break;
      case 192128: launch_fast<6, 4>(p, st); break;
      case 192064: launch_fast<6, 2>(p, st); break;
      default: done = false;
    }
    if (done) return ynb::check_launch("yn_sample_pdf_merge");
  }
  const int wpb = 8;
  const size_t smem = (size_t)wpb * (3 * P + n2) * sizeof(float);
  if (smem > 200 * 1024) return ynb::fail(YN_ERR_UNSUPPORTED, "yn_sample_pdf_merge: P + n_new too large for shared memory");
  static size_t configured = 0;
  if (configured < smem) {
    cudaFuncSetAttribute(ynb::sample_pdf_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  ynb::sample_pdf_merge_kernel<<<(unsigned)((R + wpb - 1) / wpb), wpb * 32, smem, st>>>(p);
  return ynb::check_launch("yn_sample_pdf_merge");
}

extern "C" int yn_sample_pdf_merge(const float* lengths, const float* weights, const float* u, int64_t u_row_stride,
                                   const int64_t* rng_state, int rng_site, float* new_lengths, int64_t* inds,
                                   int32_t* flag, int64_t R, int P, int n_new, int add_input_samples, void* stream) {
  return launch_pdf(lengths, weights, u, u_row_stride, rng_state, rng_site, new_lengths, inds, flag, R, P, n_new,
                    add_input_samples, 0, 1, stream);
}

extern "C" int yn_sample_pdf(const float* bins, const float* weights, const float* u, int64_t u_row_stride,
                             float* samples, int64_t* inds, int32_t* flag, int64_t R, int n_bins, int n_samples,
                             void* stream) {
  return launch_pdf(bins, weights, u, u_row_stride, nullptr, 0, samples, inds, flag, R, n_bins, n_samples, 0, 1, 0, stream);
}
