// Shared definitions of the fused NeRF-MLP kernels: architecture bookkeeping, the flat fp32 parameter
// layout (state-dict order of yanerf/pipelines/models/nerf_mlp.py:61-83), the tensor-core weight image
// and the activation stash.
#pragma once
#include <stdint.h>

#include "../../include/yanerf_b200.h"

namespace ynb {

constexpr int kInner = 256;        // inner trunk width (MLPWithInputSkips default, nerf_mlp.py:225)
constexpr int kTileM = 128;        // points per tile (= UMMA M)
constexpr int kBlkBytes = 16384;   // one [128 x 64] 16-bit operand block
constexpr int kMaxLayers = 12;
constexpr int kDirPad = 128;       // padded width of the colour hidden layer
constexpr int kBiasBlkBytes = 4096;  // [128 x 16] 16-bit, no swizzle (8x8 core matrices)
constexpr int kHeadBlkBytes = 4096;  // colour head [16 x 128] 16-bit: two swizzled [16 x 64] sub-blocks
constexpr int kDensBlkBytes = 8192;  // density head [16 x 256] 16-bit: four swizzled [16 x 64] sub-blocks
constexpr int kMaskBytes = 4096;     // sign mask of one layer and tile: 128 rows x 256 bits
constexpr int kHeadN = 16;           // UMMA N of the colour head (color_dim <= 3 rows used)

struct Arch {
  int n_layers;
  uint32_t skip_mask;
  int n_freq_xyz, n_freq_dir;
  int hidden_last, hidden_dir, color_dim;
  int fmt;

  __host__ __device__ int embed_xyz() const { return 3 * (2 * n_freq_xyz + 1); }
  __host__ __device__ int embed_dir() const { return 3 * (2 * n_freq_dir + 1); }
  // mma layer numbering: l < n_layers trunk, l == n_layers the intermediate layer, l == n_layers + 1 the colour hidden
  // layer.  The intermediate layer is linear and feeds only the colour hidden layer, so the kernels never execute it: the
  // weight image of the colour hidden layer holds the PRODUCT  W_ci = W_c[:, :H] W_i  (fp32, rounded once) and its bias
  // stage  b_ci = W_c[:, :H] b_i;  `merged(l)` marks the skipped layer.  Parameters and gradients stay per layer.
  __host__ __device__ int n_mma_layers() const { return n_layers + 2; }
  __host__ __device__ bool merged(int l) const { return l == n_layers; }
  __host__ __device__ bool has_emb(int l) const { return l == 0 || (l < n_layers && ((skip_mask >> l) & 1u)); }
  __host__ __device__ int nkb_hidden(int l) const { return l == 0 ? 0 : 4; }
  __host__ __device__ int nkb(int l) const { return nkb_hidden(l) + (has_emb(l) ? 1 : 0); }
  __host__ __device__ int nnh(int l) const { return l == n_layers + 1 ? 1 : 2; }
  // trunk / intermediate layers without an embedding K-block fold their bias in through one extra 4 KB
  // "bias block": a K=16 MMA of the embedding block's last slice (whose channel 63 is the constant 1) with a
  // [128 x 16] weight slice holding the bias in column 15.  Layers WITH an embedding block carry the bias in
  // column 63 of that block.  The colour hidden layer adds its per-ray bias in the epilogue.
  __host__ __device__ bool has_bias_stage(int l) const { return !has_emb(l); }  // (colour hidden: b_ci)
  __host__ __device__ int stages_per_half(int l) const { return nkb(l) + (has_bias_stage(l) ? 1 : 0); }
  __host__ __device__ int stages(int l) const { return merged(l) ? 0 : stages_per_half(l) * nnh(l); }
  __host__ __device__ int stage_offset(int l) const {
    int s = 0;
    for (int i = 0; i < l; ++i) s += stages(i);
    return s;
  }
  // after the layers: one 4 KB stage with the colour head, W2 as a [16 x 128] K-major operand (two [16 x 64]
  // swizzled sub-blocks at byte offsets 0 and 2048; rows >= color_dim are zero)
  // (before it: one 8 KB stage with the density head, w_d as row 0 of a [16 x 256] K-major operand, four [16 x 64]
  // sub-blocks 2 KB apart: it is multiplied against the same activations as the colour hidden layer, into 16 accumulator
  // columns that layer leaves free)
  __host__ __device__ int density_stage() const { return stage_offset(n_mma_layers()); }
  __host__ __device__ int head_stage() const { return density_stage() + 1; }
  __host__ __device__ int total_stages() const { return head_stage() + 1; }

  // true input / output widths of mma layer l
  __host__ __device__ int hidden_in(int l) const {
    if (l == 0) return 0;
    if (l <= n_layers - 1) return kInner;
    return hidden_last;  // intermediate and colour hidden read the last trunk output
  }
  __host__ __device__ int dout(int l) const {
    if (l < n_layers - 1) return kInner;
    if (l == n_layers - 1 || l == n_layers) return hidden_last;
    return hidden_dir;
  }
  __host__ __device__ int din(int l) const {  // columns of the fp32 weight matrix
    if (l < n_layers) return hidden_in(l) + (has_emb(l) ? embed_xyz() : 0);
    if (l == n_layers) return hidden_last;
    return hidden_last + embed_dir();
  }
  // offsets (in floats) inside the flat parameter vector
  __host__ __device__ int64_t w_offset(int l) const {
    int64_t o = 0;
    for (int i = 0; i < l && i < n_layers; ++i) o += (int64_t)dout(i) * din(i) + dout(i);
    if (l <= n_layers) return o;                                   // trunk l or intermediate
    o += (int64_t)hidden_last * hidden_last + hidden_last;          // intermediate
    o += hidden_last + 1;                                           // density
    return o;                                                       // colour hidden
  }
  __host__ __device__ int64_t b_offset(int l) const { return w_offset(l) + (int64_t)dout(l) * din(l); }
  __host__ __device__ int64_t density_w_offset() const {
    return w_offset(n_layers) + (int64_t)hidden_last * hidden_last + hidden_last;
  }
  __host__ __device__ int64_t density_b_offset() const { return density_w_offset() + hidden_last; }
  __host__ __device__ int64_t color2_w_offset() const { return b_offset(n_layers + 1) + hidden_dir; }
  __host__ __device__ int64_t color2_b_offset() const { return color2_w_offset() + (int64_t)color_dim * hidden_dir; }
  __host__ __device__ int64_t param_count() const { return color2_b_offset() + color_dim; }

  // aux fp32 buffer: padded biases of the n_layers + 1 wide layers, then the small heads
  __host__ __device__ int aux_bias(int l) const { return l * kInner; }
  __host__ __device__ int aux_wd() const { return (n_layers + 1) * kInner; }
  __host__ __device__ int aux_bd() const { return aux_wd() + kInner; }
  __host__ __device__ int aux_w2() const { return aux_bd() + 4; }
  __host__ __device__ int aux_b2() const { return aux_w2() + 4 * kDirPad; }
  __host__ __device__ int aux_wci() const { return aux_b2() + 4; }                 // [kDirPad][kInner] fp32: W_c[:, :H] W_i
  __host__ __device__ int aux_bci() const { return aux_wci() + kDirPad * kInner; }  // [kDirPad]: W_c[:, :H] b_i
  __host__ __device__ int aux_floats() const { return aux_bci() + kDirPad; }

  // weight image: forward stages (W as [n][k] K-major blocks), then for the backward data-gradient the
  // transposed stages (W^T as [k][n] blocks): see mlp_pack.cu
  __host__ __device__ int bwd_stages(int l) const {
    // dgrad of layer l: output columns = hidden_in(l) padded to 256 (none for layer 0), reduction over dout (256 / 128)
    if (l == 0 || merged(l)) return 0;
    return 2 * (l == n_layers + 1 ? 2 : 4);
  }
  __host__ __device__ int bwd_stage_offset(int l) const {
    int s = total_stages();
    for (int i = n_mma_layers() - 1; i > l; --i) s += bwd_stages(i);
    return s;
  }
  __host__ __device__ int total_stages_all() const { return bwd_stage_offset(0) + bwd_stages(0); }

  // stash per tile: embedding block, then the 16-bit output image of every mma layer (4 blocks; 2 for colour hidden),
  // then the ReLU sign masks: one bit per activation, [mask][128 rows][8 words] = 4 KB per ReLU layer (mask m = trunk
  // layer m for m < n_layers, m = n_layers for the colour hidden layer); bit (31 - j) of word q = "pre-activation of
  // column 32 q + j is negative".  The data-gradient kernel reads these 36 KB per tile instead of 608 KB of activations.
  __host__ __device__ int act_blocks_per_tile() const { return 1 + 4 * (n_layers + 1) + 2; }
  __host__ __device__ int mask_blocks_per_tile() const { return ((n_layers + 1) * kMaskBytes + kBlkBytes - 1) / kBlkBytes; }
  __host__ __device__ int stash_blocks_per_tile() const { return act_blocks_per_tile() + mask_blocks_per_tile(); }
  __host__ __device__ int stash_block_of_layer(int l) const { return 1 + 4 * l; }
  __host__ __device__ size_t mask_offset(int m) const { return (size_t)act_blocks_per_tile() * kBlkBytes + (size_t)m * kMaskBytes; }
};

inline Arch arch_from_c(const yn_mlp_arch* a) {
  Arch r;
  r.n_layers = a->n_layers;
  r.skip_mask = a->skip_mask;
  r.n_freq_xyz = a->n_freq_xyz;
  r.n_freq_dir = a->n_freq_dir;
  r.hidden_last = a->hidden_last;
  r.hidden_dir = a->hidden_dir;
  r.color_dim = a->color_dim;
  r.fmt = a->fmt;
  return r;
}

// sets the thread-local error string; returns code
int fail(int code, const char* fmt, ...);
int check_arch(const yn_mlp_arch* a);
int check_launch(const char* what);
bool first_use(const void* kernel);  // true the first time a kernel pointer is seen (one-time attribute setup)

}  // namespace ynb
