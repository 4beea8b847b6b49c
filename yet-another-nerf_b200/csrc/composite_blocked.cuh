// Lane-blocked compositing kernels for P = 32*S samples per ray (S = 2, 4, 6: the lego / fern shapes) and C = 3.
// Lane i of the ray's warp owns the S CONTIGUOUS samples [i*S, i*S + S): the transmittance scan becomes S serial
// fp64 adds per lane plus ONE warp scan of the lane totals (instead of one warp scan per 32 samples), every
// per-sample quantity stays in registers (the backward needs no second sweep), and all global accesses are 8-byte
// vectors whose union over the warp is the ray's contiguous row.
#pragma once

namespace ynb {

template <int S>
__device__ __forceinline__ void load_row(const float* __restrict__ src, float (&dst)[S]) {
#pragma unroll
  for (int k = 0; k < S / 2; ++k) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(src) + k);
    dst[2 * k] = v.x;
    dst[2 * k + 1] = v.y;
  }
}
template <int S>
__device__ __forceinline__ void store_row(float* __restrict__ dst, const float (&src)[S]) {
#pragma unroll
  for (int k = 0; k < S / 2; ++k) reinterpret_cast<float2*>(dst)[k] = make_float2(src[2 * k], src[2 * k + 1]);
}

__device__ __forceinline__ double warp_scan_incl_up(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
__device__ __forceinline__ double warp_scan_incl_down(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_down_sync(0xffffffffu, v, o);
    if (lane + o < 32) v += t;
  }
  return v;
}

template <int S>
struct RayState {
  float z[S], sr[S], delta[S], x[S], E[S], Eprev[S], ex[S], T[S], alpha[S], w[S];
  float E_last;
};

// forward quantities of the lane's S samples (shared by the forward and the backward kernel)
// kRng: the density noise is drawn in the kernel (training with a device generator); a separate instantiation so that
// the plain kernels keep their register budget
template <int S, bool kRng>
__device__ __forceinline__ void march_blocked(const MarchParams& p, int64_t base, int lane, float dn, RayState<S>& s) {
  const int s0 = lane * S;
  load_row<S>(p.z + base + s0, s.z);
  load_row<S>(p.sigma + base + s0, s.sr);
  if (p.cfg.density_noise_std > 0.f && (kRng || p.noise != nullptr)) {
    float nz[S];
    if (!kRng) {
      load_row<S>(p.noise + base + s0, nz);
    } else {  // in-kernel draws: the same (ray, sample) -> value map in the forward and the backward kernel
      NormalRow gen(p.rng, base / (32 * S));
#pragma unroll
      for (int k = 0; k < S; ++k) nz[k] = gen.get(s0 + k);
    }
#pragma unroll
    for (int k = 0; k < S; ++k) s.sr[k] = s.sr[k] + nz[k] * p.cfg.density_noise_std;
  }
  const float z_next_lane = __shfl_down_sync(0xffffffffu, s.z[0], 1);
  double run = 0.0;
  double pre[S];
#pragma unroll
  for (int k = 0; k < S; ++k) {
    const float zn = (k + 1 < S) ? s.z[k + 1 < S ? k + 1 : k] : z_next_lane;
    const bool last = (lane == 31 && k == S - 1);
    float delta = last ? p.cfg.background_opacity : (zn - s.z[k]);
    delta = delta * dn;
    const float dens = fmaxf(s.sr[k], 0.f) + p.cfg.background_density_bias;
    s.delta[k] = delta;
    s.x[k] = delta * dens;
    run += static_cast<double>(s.x[k]);
    pre[k] = run;
  }
  const double incl = warp_scan_incl_up(run, lane);
  double excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) excl = 0.0;
#pragma unroll
  for (int k = 0; k < S; ++k) s.E[k] = expf(-static_cast<float>(excl + pre[k]));
  float Ein = __shfl_up_sync(0xffffffffu, s.E[S - 1], 1);
#pragma unroll
  for (int k = 0; k < S; ++k) {
    s.Eprev[k] = k == 0 ? Ein : s.E[k - 1 >= 0 ? k - 1 : 0];
    s.T[k] = (lane == 0 && k == 0) ? 1.f : 1.f - (1.f - s.Eprev[k]);
    s.ex[k] = expf(-s.x[k]);
    s.alpha[k] = 1.f - s.ex[k];
    s.w[k] = s.alpha[k] * s.T[k];
  }
  s.E_last = __shfl_sync(0xffffffffu, s.E[S - 1], 31);
}

template <int S, bool kRng>
__global__ void __launch_bounds__(256) composite_fwd_blocked_kernel(const MarchParams p) {
  constexpr int C = 3;
  const int lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ray >= p.R) return;
  const int64_t base = ray * (32 * S);
  const float dx = __ldg(p.dirs + ray * 3), dy = __ldg(p.dirs + ray * 3 + 1), dz = __ldg(p.dirs + ray * 3 + 2);
  const float dn = sqrtf(dx * dx + dy * dy + dz * dz);
  float bgv[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int cc = p.cfg.bg_channels == 1 ? 0 : c;
    bgv[c] = p.bg ? __ldg(p.bg + ray * p.cfg.bg_channels + cc) : p.cfg.bg_const[cc];
  }
  float col[S * C];
  load_row<S * C>(p.rgb + (base + lane * S) * C, col);
  RayState<S> s;
  march_blocked<S, kRng>(p, base, lane, dn, s);
  store_row<S>(p.weights + base + lane * S, s.w);
  float depth = 0.f, feat[C] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < S; ++k) {
    depth += s.w[k] * s.z[k];
    const bool last_hard = p.cfg.hard_background && lane == 31 && k == S - 1;
#pragma unroll
    for (int c = 0; c < C; ++c) feat[c] += s.w[k] * (last_hard ? bgv[c] : col[k * C + c]);
  }
  depth = warp_sum(depth);
#pragma unroll
  for (int c = 0; c < C; ++c) feat[c] = warp_sum(feat[c]);
  const float opacity = 1.f - s.E_last;
  if (lane == 0) {
    p.depths[ray] = depth;
    p.opacities[ray] = opacity;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float f = feat[c];
      if (!p.cfg.hard_background) {
        const float a = p.cfg.blend_output ? opacity : 1.f;
        f = a * f + (1.f - opacity) * bgv[c];
      }
      p.features[ray * C + c] = f;
    }
  }
}

template <int S, bool kRng>
__global__ void __launch_bounds__(256) composite_bwd_blocked_kernel(const MarchParams p) {
  constexpr int C = 3;
  const int lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ray >= p.R) return;
  const int64_t base = ray * (32 * S);
  const float dx = __ldg(p.dirs + ray * 3), dy = __ldg(p.dirs + ray * 3 + 1), dz = __ldg(p.dirs + ray * 3 + 2);
  const float dn = sqrtf(dx * dx + dy * dy + dz * dz);
  float bgv[C], df[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int cc = p.cfg.bg_channels == 1 ? 0 : c;
    bgv[c] = p.bg ? __ldg(p.bg + ray * p.cfg.bg_channels + cc) : p.cfg.bg_const[cc];
    df[c] = __ldg(p.d_features + ray * C + c);
  }
  const float ddepth = p.d_depths ? __ldg(p.d_depths + ray) : 0.f;
  const float dopac = p.d_opacities ? __ldg(p.d_opacities + ray) : 0.f;
  const bool hard = p.cfg.hard_background != 0;
  const bool blend = p.cfg.blend_output != 0 && !hard;
  float col[S * C];
  load_row<S * C>(p.rgb + (base + lane * S) * C, col);
  float dwt[S];
#pragma unroll
  for (int k = 0; k < S; ++k) dwt[k] = 0.f;
  if (p.d_weights) load_row<S>(p.d_weights + base + lane * S, dwt);
  RayState<S> s;
  march_blocked<S, kRng>(p, base, lane, dn, s);
  const float opacity = 1.f - s.E_last;
  const float a = blend ? opacity : 1.f;
  float g_op = dopac;
  if (!hard) {
    float feat[C] = {0.f, 0.f, 0.f};
    if (blend) {
#pragma unroll
      for (int k = 0; k < S; ++k)
#pragma unroll
        for (int c = 0; c < C; ++c) feat[c] += s.w[k] * col[k * C + c];
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float fc = blend ? warp_sum(feat[c]) : 0.f;
      g_op += df[c] * (blend ? (fc - bgv[c]) : -bgv[c]);
    }
  }
  // g_k = dL/dw_k; contribution of sample i to every earlier x_k is g_i alpha_i E_{i-1} (none for the first sample)
  float g[S];
  double contrib[S];
  double lane_total = 0.0;
#pragma unroll
  for (int k = 0; k < S; ++k) {
    const bool last_hard = hard && lane == 31 && k == S - 1;
    float gk = ddepth * s.z[k] + dwt[k];
#pragma unroll
    for (int c = 0; c < C; ++c) gk += a * df[c] * (last_hard ? bgv[c] : col[k * C + c]);
    g[k] = gk;
    contrib[k] = (lane == 0 && k == 0) ? 0.0 : static_cast<double>(gk * s.alpha[k] * s.Eprev[k]);
    lane_total += contrib[k];
  }
  // exclusive suffix over lanes; exactly zero for the last lane
  const double after = warp_scan_incl_down(lane_total, lane) - lane_total;
  float dsig[S], dcol[S * C];
  double suf = 0.0;  // contributions of later samples of this lane
#pragma unroll
  for (int k = S - 1; k >= 0; --k) {
    const bool last = lane == 31 && k == S - 1;
    const float suffix = last ? 0.f : static_cast<float>(suf + after);
    const float dLdx = g[k] * s.T[k] * s.ex[k] - suffix + g_op * s.E_last;
    const float d = dLdx * s.delta[k];
    dsig[k] = s.sr[k] > 0.f ? d : 0.f;
    const bool last_hard = hard && last;
#pragma unroll
    for (int c = 0; c < C; ++c) dcol[k * C + c] = last_hard ? 0.f : a * df[c] * s.w[k];
    suf += contrib[k];
  }
  store_row<S>(p.d_sigma + base + lane * S, dsig);
  store_row<S * C>(p.d_rgb + (base + lane * S) * C, dcol);
}

}  // namespace ynb
