// Weight preparation for the fused NeRF-MLP kernels:
//   yn_mlp_pack_weights  fp32 master parameters (state-dict order, nerf_mlp.py:61-83) -> 16-bit tensor-core
//                        blocks [128 x 64] in the 128-byte-swizzled shared-memory image the UMMA descriptors
//                        expect, in the exact order the kernels stream them, plus a small fp32 "aux" buffer
//                        (padded biases, density head, colour head).
//   yn_mlp_dirbias       the per-ray half of LinearWithRepeat (models/utils.py:207-211) on the harmonic
//                        embedding of the normalised directions (nerf_mlp.py:97-115).
#include <cuda_runtime.h>

#include "mlp_common.cuh"
#include "sm100_ptx.cuh"

namespace ynb {

// W_ci = W_c[:, :H] W_i and b_ci = W_c[:, :H] b_i (fp32) into the aux buffer: the merged intermediate + colour hidden layer
// (mlp_common.cuh).  Block jt owns one row of W_c (kept in shared memory); thread (part, k) owns column k and a quarter of
// the reduction, the four partial sums are folded in a fixed order: deterministic (the same weights always give the same
// image), coalesced reads of W_i, and a dependent-FMA chain of 64 instead of 256 (the kernel is pure latency: 37 -> ~12 us).
constexpr int kFuseParts = 4;
__global__ void __launch_bounds__(kFuseParts * 256) fuse_color_kernel(const Arch A, const float* __restrict__ params, float* __restrict__ aux) {
  const int jt = blockIdx.x;
  const int n = A.n_layers, H = A.hidden_last, dinc = A.din(n + 1);
  const float* Wi = params + A.w_offset(n);
  const float* bi = params + A.b_offset(n);
  const float* Wc = params + A.w_offset(n + 1);
  __shared__ float s_wc[kInner];
  __shared__ float s_part[kFuseParts][kInner];
  const int t = threadIdx.x & 255, part = threadIdx.x >> 8;
  if (part == 0) s_wc[t] = (jt < A.hidden_dir && t < H) ? Wc[(int64_t)jt * dinc + t] : 0.f;
  __syncthreads();
  float acc = 0.f;
  if (t < H) {
    const int o0 = part * (kInner / kFuseParts), o1 = min(o0 + kInner / kFuseParts, H);
#pragma unroll 8
    for (int o = o0; o < o1; ++o) acc = fmaf(s_wc[o], Wi[(int64_t)o * H + t], acc);
  }
  s_part[part][t] = acc;
  __syncthreads();
  if (part == 0) {
    float v = s_part[0][t];
#pragma unroll
    for (int q = 1; q < kFuseParts; ++q) v += s_part[q][t];
    aux[A.aux_wci() + jt * kInner + t] = v;
  }
  if (threadIdx.x == 256) {  // (a thread of the second part: the bias product runs beside the fold)
    float b = 0.f;
    for (int o = 0; o < H; ++o) b = fmaf(s_wc[o], bi[o], b);
    aux[A.aux_bci() + jt] = b;
  }
}

template <int kFmt>
__global__ void pack_weights_kernel(const Arch A, const float* __restrict__ params, const float* __restrict__ aux,
                                    uint8_t* __restrict__ wpack, int n_units) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_units) return;
  const int s = idx >> 10;  // 1024 16-byte units per block
  const int unit = idx & 1023;
  const int r = unit >> 3, u = unit & 7;
  const int L = A.n_mma_layers();
  const int fwd_total = A.total_stages();
  float vals[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) vals[i] = 0.f;

  if (s < fwd_total) {
    // forward block (l, nh, kb): rows = output features nh*128 + r, cols = input columns of K-block kb
    int l = 0, off = 0;
    while (l < L && s >= off + A.stages(l)) { off += A.stages(l); ++l; }
    if (l == L) {
      // density head stage: w_d as row 0 of a [16 x 256] K-major operand (4 sub-blocks); colour head stage: W2
      // [color_dim x hidden_dir] as a [16 x 128] operand (2 sub-blocks); sub-block kb at byte offset kb * 2048
      const bool dens = s == A.density_stage();
      uint4 out = make_uint4(0u, 0u, 0u, 0u);
      size_t dst = (size_t)unit * 16;
      if (unit < (dens ? 512 : 256)) {
        const int kb = unit >> 7, rr = (unit >> 3) & 15, uu = unit & 7;
        float w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int k = kb * 64 + uu * 8 + i;
          if (dens) w[i] = (rr == 0 && k < A.hidden_last) ? params[A.density_w_offset() + k] : 0.f;
          else w[i] = (rr < A.color_dim && k < A.hidden_dir) ? params[A.color2_w_offset() + (int64_t)rr * A.hidden_dir + k] : 0.f;
        }
        out.x = Half2Pack<kFmt>::pack(w[0], w[1]);
        out.y = Half2Pack<kFmt>::pack(w[2], w[3]);
        out.z = Half2Pack<kFmt>::pack(w[4], w[5]);
        out.w = Half2Pack<kFmt>::pack(w[6], w[7]);
        dst = (size_t)kb * 2048 + rr * 128 + ((uu ^ (rr & 7)) << 4);
      }
      *reinterpret_cast<uint4*>(wpack + (size_t)s * kBlkBytes + dst) = out;
      return;
    }
    const int local = s - off;
    const int nkb = A.nkb(l);
    const int per_half = A.stages_per_half(l);
    const int nh = local / per_half, kb = local % per_half;
    const int din = A.din(l);
    if (kb == nkb) {
      // bias block: [128 x 16] no-swizzle image in the first 4 KB of the slot (rest zero): column 15 = bias
      uint4 out = make_uint4(0u, 0u, 0u, 0u);
      if (unit < 256) {
        const int grp = unit >> 4, kc = (unit >> 3) & 1, rr = unit & 7;
        const int n = nh * 128 + grp * 8 + rr;
        if (kc == 1 && n < A.dout(l))
          out.w = Half2Pack<kFmt>::pack(0.f, l == A.n_layers + 1 ? aux[A.aux_bci() + n] : params[A.b_offset(l) + n]);
      }
      *reinterpret_cast<uint4*>(wpack + (size_t)s * kBlkBytes + (size_t)unit * 16) = out;
      return;
    }
    const int n = nh * 128 + r;
    if (n < A.dout(l)) {
      const float* W = params + A.w_offset(l);
      const bool emb_blk = kb >= A.nkb_hidden(l);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = u * 8 + i;
        int col = -1;
        if (emb_blk) {
          if (c < A.embed_xyz()) col = A.hidden_in(l) + c;
          // channel 63 of the embedding block is the constant 1: its weight is the layer's bias
          if (c == 63 && l <= A.n_layers) vals[i] = params[A.b_offset(l) + n];
        } else {
          const int k = kb * 64 + c;
          if (k < A.hidden_in(l)) col = k;
        }
        if (col >= 0) vals[i] = l == A.n_layers + 1 ? aux[A.aux_wci() + n * kInner + col] : W[(int64_t)n * din + col];
      }
    }
  } else {
    // data-gradient block of layer l (walked in reverse layer order): rows = input feature nh*128 + r,
    // cols = output feature kb*64 + c, value W_l[out][in]
    int l = L - 1, off = fwd_total;
    while (l > 0 && s >= off + A.bwd_stages(l)) { off += A.bwd_stages(l); --l; }
    const int local = s - off;
    const int nkb = A.bwd_stages(l) / 2;
    const int nh = local / nkb, kb = local % nkb;
    const int n = nh * 128 + r;
    const int din = A.din(l);
    if (n < A.hidden_in(l)) {
      const float* W = params + A.w_offset(l);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = kb * 64 + u * 8 + i;
        if (k < A.dout(l)) vals[i] = l == A.n_layers + 1 ? aux[A.aux_wci() + k * kInner + n] : W[(int64_t)k * din + n];
      }
    }
  }
  uint4 out;
  out.x = Half2Pack<kFmt>::pack(vals[0], vals[1]);
  out.y = Half2Pack<kFmt>::pack(vals[2], vals[3]);
  out.z = Half2Pack<kFmt>::pack(vals[4], vals[5]);
  out.w = Half2Pack<kFmt>::pack(vals[6], vals[7]);
  *reinterpret_cast<uint4*>(wpack + (size_t)s * kBlkBytes + r * 128 + ((u ^ (r & 7)) << 4)) = out;
}

__global__ void pack_aux_kernel(const Arch A, const float* __restrict__ params, float* __restrict__ aux, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = 0.f;
  if (i < A.aux_wd()) {
    const int l = i / kInner, c = i % kInner;
    if (c < A.dout(l)) v = params[A.b_offset(l) + c];
  } else if (i < A.aux_bd()) {
    const int c = i - A.aux_wd();
    if (c < A.hidden_last) v = params[A.density_w_offset() + c];
  } else if (i < A.aux_w2()) {
    if (i == A.aux_bd()) v = params[A.density_b_offset()];
  } else if (i < A.aux_b2()) {
    const int c = (i - A.aux_w2()) / kDirPad, j = (i - A.aux_w2()) % kDirPad;
    if (c < A.color_dim && j < A.hidden_dir) v = params[A.color2_w_offset() + (int64_t)c * A.hidden_dir + j];
  } else {
    const int c = i - A.aux_b2();
    if (c < A.color_dim) v = params[A.color2_b_offset() + c];
  }
  aux[i] = v;
}

// one block = kRays rays x 128 outputs: the 27-channel embeddings of the rays are built cooperatively in shared
// memory, then thread j keeps its weight row in registers and walks the rays (coalesced 512-byte stores).
// kRays = 128 for full-image renders (the per-block weight-row loads and the set-up are amortised over 4x the rays),
// 32 for training batches (enough blocks for every SM).
template <int kRays>
__global__ void __launch_bounds__(128) dirbias_kernel(const Arch A, const float* __restrict__ params,
                                                     const float* __restrict__ directions,
                                                     float* __restrict__ dirbias, int64_t R) {
  __shared__ __align__(16) float s_emb[kRays][32];  // rows read back as float4 broadcasts (a scalar read per FMA was LDS-bound)
  __shared__ float s_dir[kRays][3];
  const int ed = A.embed_dir();  // <= 32 (check_arch)
  const int nf = A.n_freq_dir;
  const int din = A.din(A.n_layers + 1);
  const float* W = params + A.w_offset(A.n_layers + 1);
  const int j = threadIdx.x;
  float w[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) w[k] = (j < A.hidden_dir && k < ed) ? __ldg(W + (int64_t)j * din + A.hidden_last + k) : 0.f;
  const float bj = j < A.hidden_dir ? __ldg(params + A.b_offset(A.n_layers + 1) + j) : 0.f;
  const int64_t ray0 = (int64_t)blockIdx.x * kRays;
  if (j < kRays) {
    const int64_t ray = ray0 + j;
    float d[3] = {0.f, 0.f, 0.f};
    if (ray < R) {
      // F.normalize(d, dim=-1) = d / max(|d|, 1e-12)  (nerf_mlp.py:105)
      const float dx = directions[ray * 3], dy = directions[ray * 3 + 1], dz = directions[ray * 3 + 2];
      const float nrm = fmaxf(sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz))), 1e-12f);
      d[0] = dx / nrm; d[1] = dy / nrm; d[2] = dz / nrm;
    }
    s_dir[j][0] = d[0]; s_dir[j][1] = d[1]; s_dir[j][2] = d[2];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kRays * 32; i += blockDim.x) {
    const int r = i >> 5, k = i & 31;
    float e = 0.f;  // channels >= embed_dir are zero
    if (k < 3 * nf) e = sinf(s_dir[r][k / nf] * exp2f((float)(k % nf)));
    else if (k < 6 * nf) e = cosf(s_dir[r][(k - 3 * nf) / nf] * exp2f((float)((k - 3 * nf) % nf)));
    else if (k < ed) e = s_dir[r][k - 6 * nf];
    s_emb[r][k] = e;
  }
  __syncthreads();
  for (int r = 0; r < kRays; ++r) {
    const int64_t ray = ray0 + r;
    if (ray >= R) break;
    float acc = bj;
    const float4* e4 = reinterpret_cast<const float4*>(s_emb[r]);
#pragma unroll
    for (int k = 0; k < 8; ++k) {  // channels >= embed_dir hold zeros
      const float4 e = e4[k];
      acc = fmaf(w[4 * k], e.x, acc);
      acc = fmaf(w[4 * k + 1], e.y, acc);
      acc = fmaf(w[4 * k + 2], e.z, acc);
      acc = fmaf(w[4 * k + 3], e.w, acc);
    }
    dirbias[ray * kDirPad + j] = acc;
  }
}

}  // namespace ynb

extern "C" int yn_mlp_pack_weights(const yn_mlp_arch* arch, const float* params, void* wpack, float* aux,
                                   void* stream) {
  if (int rc = ynb::check_arch(arch)) return rc;
  if (!params || !wpack || !aux) return ynb::fail(YN_ERR_INVALID_ARGUMENT, "yn_mlp_pack_weights: null pointer");
  const ynb::Arch A = ynb::arch_from_c(arch);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n_units = A.total_stages_all() * 1024;
  ynb::fuse_color_kernel<<<ynb::kDirPad, ynb::kFuseParts * 256, 0, st>>>(A, params, aux);  // one block per row of W_c
  if (A.fmt == 1)
    ynb::pack_weights_kernel<1><<<(n_units + 255) / 256, 256, 0, st>>>(A, params, aux, static_cast<uint8_t*>(wpack), n_units);
  else
    ynb::pack_weights_kernel<0><<<(n_units + 255) / 256, 256, 0, st>>>(A, params, aux, static_cast<uint8_t*>(wpack), n_units);
  const int na = A.aux_wci();  // the small heads and padded biases; [aux_wci, aux_floats) was written by fuse_color_kernel
  ynb::pack_aux_kernel<<<(na + 255) / 256, 256, 0, st>>>(A, params, aux, na);
  return ynb::check_launch("yn_mlp_pack_weights");
}

extern "C" int yn_mlp_dirbias(const yn_mlp_arch* arch, const float* params, const float* directions,
                              float* dirbias, int64_t R, void* stream) {
  if (int rc = ynb::check_arch(arch)) return rc;
  if (R < 0) return ynb::fail(YN_ERR_INVALID_ARGUMENT, "yn_mlp_dirbias: R < 0");
  if (R == 0) return YN_OK;
  if (!params || !directions || !dirbias) return ynb::fail(YN_ERR_INVALID_ARGUMENT, "yn_mlp_dirbias: null pointer");
  const ynb::Arch A = ynb::arch_from_c(arch);
  if (R >= 64 * 1024)
    ynb::dirbias_kernel<128><<<(unsigned)((R + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(A, params, directions, dirbias, R);
  else
    ynb::dirbias_kernel<32><<<(unsigned)((R + 31) / 32), 128, 0, static_cast<cudaStream_t>(stream)>>>(A, params, directions, dirbias, R);
  return ynb::check_launch("yn_mlp_dirbias");
}

extern "C" int64_t yn_mlp_param_count(const yn_mlp_arch* arch) {
  if (ynb::check_arch(arch)) return -1;
  return ynb::arch_from_c(arch).param_count();
}
extern "C" int64_t yn_mlp_wpack_bytes(const yn_mlp_arch* arch) {
  if (ynb::check_arch(arch)) return -1;
  return (int64_t)ynb::arch_from_c(arch).total_stages_all() * ynb::kBlkBytes;
}
extern "C" int64_t yn_mlp_aux_floats(const yn_mlp_arch* arch) {
  if (ynb::check_arch(arch)) return -1;
  return ynb::arch_from_c(arch).aux_floats();
}
extern "C" int64_t yn_mlp_stash_bytes(const yn_mlp_arch* arch, int64_t n_points) {
  if (ynb::check_arch(arch) || n_points < 0) return -1;
  const int64_t n_tiles = (n_points + ynb::kTileM - 1) / ynb::kTileM;
  // the forward kernel works on tile pairs; keep room for the (masked) odd partner
  return ((n_tiles + 1) / 2 * 2) * (int64_t)ynb::arch_from_c(arch).stash_blocks_per_tile() * ynb::kBlkBytes;
}
