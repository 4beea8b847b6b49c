// Fast path of yn_sample_pdf_merge for the lego / fern shapes: P = 32*PS input depths (PS = 2, 4, 6), 32*NPL new
// samples (NPL = 2, 4), refiner mode with the input samples appended.  Same arithmetic (and therefore the same bits)
// as sample_pdf_merge_kernel; the differences are purely structural:
//   * lane-blocked layouts (lane i owns PS consecutive depths / NPL consecutive draws): 8/16-byte vector accesses;
//   * the cdf is padded with +inf to a power of two so searchsorted is a branch-free 6..8 step bisection;
//   * the new samples are sorted in registers (skipped when the draws were already ascending);
//   * the final merge is a merge-path split: every lane emits PS + NPL consecutive outputs.
#pragma once

namespace ynb {

// aten_lane_partial<8> with the row length known at compile time (rows shorter than 512 never reach the cascade
// levels, so the order collapses to: 4 interleaved accumulators over groups of 4 vectors, leftover vectors into
// accumulator 0, then accumulators 1..3 folded in)
template <int K>
__device__ __forceinline__ float aten_lane_partial8_ct(const float* wp, int j) {
  constexpr int vec_size = K / 8, size_ilp = vec_size / 4;
  static_assert(size_ilp < 16, "cascade levels not needed below 512 elements");
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < size_ilp; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[c] = __fadd_rn(acc[c], wp[(i * 4 + c) * 8 + j]);
  float part = acc[0];
#pragma unroll
  for (int v = size_ilp * 4; v < vec_size; ++v) part = __fadd_rn(part, wp[v * 8 + j]);
  part = __fadd_rn(part, acc[1]);
  part = __fadd_rn(part, acc[2]);
  part = __fadd_rn(part, acc[3]);
  return part;
}

template <int PS, int NPL, bool kRng>
__global__ void __launch_bounds__(256) sample_pdf_merge_fast_kernel(const PdfParams p) {
  constexpr int P = 32 * PS, K = P - 2, NB = P - 1, N = 32 * NPL, OPL = PS + NPL;
  constexpr int CDFN = P <= 64 ? 64 : (P <= 128 ? 128 : 256);  // padded cdf length
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  if (ray >= p.R) return;
  const int per_warp = CDFN + 2 * P + p.sort_pow2;
  float* s_cdf = smem + (size_t)wib * per_warp;  // [CDFN]: first w + eps (at [k + 1]), then the cdf, +inf padded
  float* s_bin = s_cdf + CDFN;                   // [NB]
  float* s_z = s_bin + P;                        // [P]
  float* s_new = s_z + P;                        // [sort_pow2]
  const float* zr = p.z + ray * P;
  const float* wr = p.w + ray * P;

  // ---- load (lane i owns depths PS*i .. PS*i + PS - 1)
  float zl[PS], wl[PS];
#pragma unroll
  for (int k = 0; k < PS / 2; ++k) {
    const float2 a = __ldg(reinterpret_cast<const float2*>(zr + lane * PS) + k);
    const float2 b = __ldg(reinterpret_cast<const float2*>(wr + lane * PS) + k);
    zl[2 * k] = a.x; zl[2 * k + 1] = a.y;
    wl[2 * k] = b.x; wl[2 * k + 1] = b.y;
  }
  const float z_next_lane = __shfl_down_sync(0xffffffffu, zl[0], 1);
  bool bad = false, unsorted = false;
#pragma unroll
  for (int r = 0; r < PS; ++r) {
    const int i = lane * PS + r;
    s_z[i] = zl[r];
    const float zn = (r + 1 < PS) ? zl[r + 1 < PS ? r + 1 : r] : z_next_lane;
    if (i + 1 < P) {
      s_bin[i] = __fsub_rn(zl[r], __fmul_rn(__fsub_rn(zl[r], zn), 0.5f));  // torch.lerp(z[1:], z[:-1], 0.5)
      unsorted |= zn < zl[r];
    }
    if (i >= 1 && i <= K) {
      const float wv = __fadd_rn(wl[r], 1e-5f);
      s_cdf[i] = wv;  // wp[k] lives at s_cdf[k + 1]
      bad |= wv <= 0.f;
    }
  }
  unsorted = __any_sync(0xffffffffu, unsorted);
  bad = __any_sync(0xffffffffu, bad);
  if (bad && lane == 0) atomicOr(p.flag, 1);
  __syncwarp();
  const float* wp = s_cdf + 1;

  // ---- sum in ATen order (K >= 8 here)
  float total = 0.f;
  {
    float part = 0.f;
    if (lane < 8) part = aten_lane_partial8_ct<K>(wp, lane);
#pragma unroll
    for (int k = (K / 8) * 8; k < K; ++k) total = __fadd_rn(total, wp[k]);
#pragma unroll
    for (int j = 0; j < 8; ++j) total = __fadd_rn(total, __shfl_sync(0xffffffffu, part, j));
  }

  // ---- pdf, cdf; lane i owns k = PS*i .. PS*i + PS - 1.
  // torch.cumsum on the CPU accumulates in fp64 and rounds every prefix to fp32.  As long as every pdf value is
  // >= 2^-28 (always inside the renderer: w + 1e-5 >= 1e-5, total <= P), each addend's lowest mantissa bit is
  // >= 2^-51 and every partial sum is < 2, so the fp64 sums are EXACT -- and a 64-bit fixed-point sum (2^-62
  // resolution) reproduces them bit for bit in any association order, on the integer pipe (fp64 runs at a few lanes
  // per clock on this part and dominated the kernel).  Rows with a smaller value take the sequential fp64 redo below.
  float pdf[PS];
  unsigned long long pre[PS], run = 0ull;
  bool ambiguous = false;
#pragma unroll
  for (int r = 0; r < PS; ++r) {
    const int k = lane * PS + r;
    pdf[r] = k < K ? __fdiv_rn(wp[k], total) : 0.f;
    if (k < K) {
      const bool fits = pdf[r] >= 0x1p-28f && pdf[r] < 2.f;  // (false for NaN)
      ambiguous |= !fits;
      run += fits ? __float2ull_rz(__fmul_rn(pdf[r], 0x1p62f)) : 0ull;  // exact: a power-of-two scaling
    }
    pre[r] = run;
  }
  unsigned long long incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  unsigned long long excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) excl = 0ull;
  ambiguous |= incl >= (1ull << 63);  // cannot happen for a normalised pdf; keeps the fixed point honest
  float cdfv[PS];
#pragma unroll
  for (int r = 0; r < PS; ++r) cdfv[r] = __fmul_rn(__ull2float_rn(excl + pre[r]), 0x1p-62f);  // RNE like fp64 -> fp32
  __syncwarp();  // every lane has consumed its wp[] before they are overwritten
#pragma unroll
  for (int r = 0; r < PS; ++r) {
    const int k = lane * PS + r;
    if (k < K) s_cdf[k + 1] = cdfv[r];
  }
  if (lane == 0) s_cdf[0] = 0.f;
  for (int i = NB + lane; i < CDFN; i += 32) s_cdf[i] = CUDART_INF_F;
  __syncwarp();
  if (__any_sync(0xffffffffu, ambiguous)) {
    if (lane == 0) {  // sequential fp64 redo = torch's own order (never taken inside the renderer)
      double acc = 0.0;
      for (int k = 0; k < K; ++k) {
        const float wv = __fadd_rn(__ldg(wr + k + 1), 1e-5f);
        acc += static_cast<double>(__fdiv_rn(wv, total));
        s_cdf[k + 1] = static_cast<float>(acc);
      }
    }
    __syncwarp();
  }

  // ---- inverse-CDF samples; lane i owns draws NPL*i .. NPL*i + NPL - 1
  float ul[NPL], v[NPL];
  if (!kRng) {
    const float* ur = p.u + ray * p.u_stride + lane * NPL;
#pragma unroll
    for (int k = 0; k < NPL / 2; ++k) {
      const float2 a = __ldg(reinterpret_cast<const float2*>(ur) + k);
      ul[2 * k] = a.x; ul[2 * k + 1] = a.y;
    }
  } else {  // in-kernel draws (training): no [R, N] tensor of uniforms in HBM
    UniformRow gen(p.rng, ray);
#pragma unroll
    for (int r = 0; r < NPL; ++r) ul[r] = gen.get(lane * NPL + r);
  }
#pragma unroll
  for (int r = 0; r < NPL; ++r) {
    const float u = ul[r];
    int pos = 0;  // number of cdf entries <= u  (searchsorted right=True)
#pragma unroll
    for (int step = CDFN / 2; step >= 1; step >>= 1)
      if (s_cdf[pos + step - 1] <= u) pos += step;
    const int ind = pos;
    const int below = ind - 1 > 0 ? ind - 1 : 0;
    const int above = ind < NB - 1 ? ind : NB - 1;
    const float cb = s_cdf[below], ca = s_cdf[above];
    const float bb = s_bin[below], ba = s_bin[above];
    float den = __fsub_rn(ca, cb);
    if (den < 1e-5f) den = 1.f;
    const float t = __fdiv_rn(__fsub_rn(u, cb), den);
    v[r] = __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));
    if (p.inds) p.inds[ray * N + lane * NPL + r] = ind;
  }

  float* outr = p.out + ray * (int64_t)(N + P);
  if (unsorted) {
    // input depths not ascending (never the case inside the renderer): sort everything together
#pragma unroll
    for (int r = 0; r < NPL; ++r) s_new[lane * NPL + r] = v[r];
    for (int i = lane; i < P; i += 32) s_new[N + i] = s_z[i];
    int n2 = 1;
    while (n2 < N + P) n2 <<= 1;
    for (int i = N + P + lane; i < n2; i += 32) s_new[i] = CUDART_INF_F;
    __syncwarp();
    warp_bitonic_sort(s_new, n2, lane);
    for (int i = lane; i < N + P; i += 32) outr[i] = s_new[i];
    return;
  }
  // ---- sort the new samples in registers unless they are already ascending
  {
    bool ok = true;
#pragma unroll
    for (int r = 0; r + 1 < NPL; ++r) ok &= v[r] <= v[r + 1];
    const float nxt = __shfl_down_sync(0xffffffffu, v[0], 1);
    ok &= (lane == 31) || (v[NPL - 1] <= nxt);
    if (!__all_sync(0xffffffffu, ok)) warp_bitonic_sort_regs<NPL>(v, lane);
  }
#pragma unroll
  for (int r = 0; r < NPL; ++r) s_new[lane * NPL + r] = v[r];
  __syncwarp();

  // ---- merge path: lane emits outputs [OPL*lane, OPL*lane + OPL); input depths first on ties
  const int d = OPL * lane;
  int lo = d - N > 0 ? d - N : 0, hi = d < P ? d : P;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (s_z[mid] <= s_new[d - 1 - mid]) lo = mid + 1; else hi = mid;
  }
  int a = lo, b = d - lo;
  float o[OPL];
#pragma unroll
  for (int k = 0; k < OPL; ++k) {
    const float za = a < P ? s_z[a] : CUDART_INF_F;
    const float nb = b < N ? s_new[b] : CUDART_INF_F;
    const bool take_a = (b >= N) || (a < P && za <= nb);
    o[k] = take_a ? za : nb;
    a += take_a ? 1 : 0;
    b += take_a ? 0 : 1;
  }
#pragma unroll
  for (int k = 0; k < OPL / 2; ++k) reinterpret_cast<float2*>(outr + d)[k] = make_float2(o[2 * k], o[2 * k + 1]);
}

}  // namespace ynb
