/*
 * yanerf_b200 -- C ABI of the B200 (sm_100a) kernels behind the yanerf.pipelines hot path.
 *
 * The reference (xk-huang/yet-another-nerf) is pure Python/PyTorch and has no FFI of its own; the
 * entry points below are the operator boundary its torch code would bind if the per-ray hot path were
 * native.  Each one names the reference function (file:line under /root/reference) it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (fp32 unless stated), rows contiguous,
 *     rays flattened to R = B*n (or B*H*W) with the sample axis innermost;
 *   - the caller allocates every output and workspace; kernels never allocate, never synchronise and
 *     run on the given stream (a cudaStream_t passed as void*);
 *   - return value: 0 on success, negative yn_status on failure; yn_last_error_string() describes the
 *     last failure of the calling thread;
 *   - optional pointers may be NULL where noted.
 */
#ifndef YANERF_B200_H_
#define YANERF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum yn_status {
  YN_OK = 0,
  YN_ERR_INVALID_ARGUMENT = -1, /* maps to ValueError on the Python side */
  YN_ERR_UNSUPPORTED = -2,      /* architecture outside the kernel family (NotImplementedError) */
  YN_ERR_CUDA = -3,             /* a CUDA runtime call / launch failed (RuntimeError) */
  YN_ERR_NO_DEVICE = -4         /* no sm_100 device */
} yn_status;

int yn_version(void);                   /* ABI version, currently 1 */
const char* yn_last_error_string(void); /* thread-local, never NULL */

/* ------------------------------------------------------------------------------------------------
 * Ray sampler: pixel coordinates -> ray bundle, depths, stratified jitter.
 * Replaces _xy_to_ray_bundle (yanerf/pipelines/ray_samplers/ray_sampler.py:249-314) and
 * _jiggle_within_stratas (ray_sampler.py:361-386).
 *   poses   [B,3,4] camera-to-world (row stride pose_row_stride floats, batch stride pose_batch_stride)
 *   focal   [B]
 *   xy      [B,n,2] float pixel coordinates, or NULL with full_grid != 0 (x = i % W, y = i / W)
 *   depths  [P] the linspace(min_depth, max_depth, P) row (ray_sampler.py:285-291)
 *   u       [B,n,P] uniform draws for the stratified jitter, or NULL for the plain depths
 *   out: origins [B,n,3], directions [B,n,3] (NOT normalised), lengths [B,n,P], xy_out [B,n,2] or NULL
 * ---------------------------------------------------------------------------------------------- */
int yn_ray_bundle(const float* poses, int64_t pose_batch_stride, int64_t pose_row_stride, const float* focal,
                  const float* xy, const float* depths, const float* u, float* origins, float* directions,
                  float* lengths, float* xy_out, int64_t B, int64_t n, int P, int width, int height, int full_grid,
                  void* stream);

/* Training pixel pick (ray_sampler.py:187-229, unmasked case: torch.multinomial(ones(H*W), n, replacement=False)):
 * n distinct pixels per image, uniformly at random, by a keyed Feistel permutation with cycle walking.
 *   seed  int64[1] in DEVICE memory (so a captured CUDA graph can be replayed with a bumped seed)
 *   out: idx int64 [B,n] flat pixel indices (x + W*y), xy float [B,n,2] or NULL */
int yn_sample_pixels(const int64_t* seed, int64_t* idx, float* xy, int64_t B, int64_t n, int width, int height,
                     void* stream);

/* ------------------------------------------------------------------------------------------------
 * NeRF MLP (yanerf/pipelines/models/nerf_mlp.py:12-289, models/utils.py:17-245).
 * Architecture family: inner trunk width 256 (the reference never forwards hidden_dim, nerf_mlp.py:88-95),
 * n_layers <= 12, skips anywhere but layer 0, 3*(2*n_freq_xyz+1) <= 64, hidden_last <= 256,
 * hidden_dir <= 128, color_dim <= 3, 3*(2*n_freq_dir+1) <= 32, latent_dim == 0.
 * ---------------------------------------------------------------------------------------------- */
typedef struct yn_mlp_arch {
  int32_t n_layers;    /* trunk layers (lego: 8) */
  uint32_t skip_mask;  /* bit l set: input of trunk layer l is cat(hidden, embedding) (lego: 1<<5) */
  int32_t n_freq_xyz;  /* lego: 10 */
  int32_t n_freq_dir;  /* lego: 4 */
  int32_t hidden_last; /* n_hidden_neurons_xyz, lego: 256 */
  int32_t hidden_dir;  /* n_hidden_neurons_dir, lego: 128 */
  int32_t color_dim;   /* lego: 3 */
  int32_t fmt;         /* tensor-core operand type: 0 = fp16, 1 = bf16 (fp32 accumulate either way) */
} yn_mlp_arch;

/* Flat fp32 parameter vector = the module's state_dict tensors concatenated in registration order
 * (xyz_encoder.mlp.{l}.0.{weight,bias}, intermediate_linear, density_layer, color_layer.0, color_layer.2). */
int64_t yn_mlp_param_count(const yn_mlp_arch* arch);
int64_t yn_mlp_wpack_bytes(const yn_mlp_arch* arch); /* tensor-core weight image (forward + backward) */
int64_t yn_mlp_aux_floats(const yn_mlp_arch* arch);  /* fp32: padded biases, head weights/biases, the product W_c[:, :H] W_i */
int64_t yn_mlp_stash_bytes(const yn_mlp_arch* arch, int64_t n_points); /* 16-bit activations + ReLU sign masks kept for backward */

/* fp32 master weights -> 16-bit swizzled tensor-core images + fp32 aux; run after every weight update.  The image is not a
 * per-layer copy: the linear intermediate layer is multiplied into the colour hidden layer (W_c[:, :H] W_i, rounded once). */
int yn_mlp_pack_weights(const yn_mlp_arch* arch, const float* params, void* wpack, float* aux, void* stream);

/* per-ray part of LinearWithRepeat (models/utils.py:207-211) + harmonic embedding of normalised
 * directions (nerf_mlp.py:97-115): dirbias[R,128] = W_c[:, H:] * emb(d/|d|) + b_c */
int yn_mlp_dirbias(const yn_mlp_arch* arch, const float* params, const float* directions, float* dirbias,
                   int64_t R, void* stream);

/* NeRFMLP.forward (nerf_mlp.py:117-177) fused with ray_bundle_to_ray_points and HarmonicEmbedding
 * (models/utils.py:90-103,214-245): origins/directions [R,3], lengths [R,P] ->
 * raw density [R,P], rgb [R,P,color_dim].  stash may be NULL (inference). */
int yn_mlp_fwd(const yn_mlp_arch* arch, const float* origins, const float* directions, const float* lengths,
               const float* dirbias, const void* wpack, const float* aux, float* density, float* rgb,
               void* stash, int64_t R, int P, void* stream);

/* Backward of yn_mlp_fwd (autograd of nerf_mlp.py:117-177): consumes d_density [R,P], d_rgb [R,P,C] and
 * `rgb` (the forward output) and the stash, ACCUMULATES into grads (flat fp32, same layout as params).  workspace: yn_mlp_bwd_workspace_bytes. */
int64_t yn_mlp_bwd_workspace_bytes(const yn_mlp_arch* arch, int64_t n_points);
int yn_mlp_bwd(const yn_mlp_arch* arch, const float* directions, const float* rgb, const float* d_density,
               const float* d_rgb, const float* params, const void* wpack, const float* aux, const void* stash,
               void* workspace, float* grads, int64_t R, int P, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Emission-absorption raymarcher (renderers/multipass_emission_absorpsion_renderer.py:154-239).
 *   raw_density [R,P], rgb [R,P,C], lengths [R,P], directions [R,3]
 *   noise [R,P] standard normal or NULL (density_noise_std == 0)
 *   bg [R,bg_channels] per-ray background or NULL -> bg_const[bg_channels] (host values)
 *   out: features [R,C], depths [R], opacities [R], weights [R,P]
 * ---------------------------------------------------------------------------------------------- */
typedef struct yn_march_cfg {
  float background_opacity;      /* 1e10 */
  float background_density_bias; /* lego: 1e-6 */
  float density_noise_std;       /* 0 in evaluation */
  int32_t blend_output;
  int32_t hard_background;
  int32_t bg_channels; /* 1 or C */
  float bg_const[4];
} yn_march_cfg;

/* noise: [R,P] N(0,1) draws (the reference's randn_like, lines 203-207) or NULL; with noise == NULL, rng_state != NULL and
 * cfg->density_noise_std > 0 the draws are generated in the kernel (see "in-kernel draws" below), identically in
 * yn_composite_fwd and yn_composite_bwd of the same step. */
int yn_composite_fwd(const yn_march_cfg* cfg, const float* raw_density, const float* rgb, const float* lengths,
                     const float* directions, const float* noise, const int64_t* rng_state, int rng_site,
                     const float* bg, float* features, float* depths, float* opacities, float* weights, int64_t R,
                     int P, int C, void* stream);

/* analytic backward (autograd of the same lines); d_depths / d_opacities / d_weights may be NULL */
int yn_composite_bwd(const yn_march_cfg* cfg, const float* raw_density, const float* rgb, const float* lengths,
                     const float* directions, const float* noise, const int64_t* rng_state, int rng_site,
                     const float* bg, const float* d_features,
                     const float* d_depths, const float* d_opacities, const float* d_weights,
                     float* d_raw_density, float* d_rgb, int64_t R, int P, int C, void* stream);

/* ------------------------------------------------------------------------------------------------
 * RayPointRefiner + sample_pdf (renderers/utils.py:36-158): midpoints, PDF->CDF, searchsorted,
 * inverse-CDF samples, concatenation with the input depths and ascending sort, one launch.
 *   lengths [R,P], weights [R,P] (raymarcher weights; the kernel uses weights[:,1:-1])
 *   u: the draws, row stride u_row_stride floats: n_new (or more) for per-ray torch.rand draws, 0 for one
 *      shared row (the deterministic linspace(0,1,n_new) of renderers/utils.py:130-132); NULL with rng_state != NULL:
 *      per-ray U[0,1) draws generated in the kernel (no [R,n_new] tensor in HBM)
 *   out: new_lengths [R, n_new + (add_input_samples ? P : 0)] sorted ascending
 *        inds [R,n_new] int64 searchsorted indices (may be NULL)
 *        flag[0] (int32, device) is set to 1 if any weight + eps <= 0 (reference raises ValueError,
 *        renderers/utils.py:123-124); the caller decides when to read it.
 * Summation orders replicate torch-CPU fp32 (sum: 8-lane x 4-accumulator cascade; cumsum: fp64 running
 * sum rounded per prefix) so indices are bit-identical to the reference's --device cpu path.
 * ---------------------------------------------------------------------------------------------- */
int yn_sample_pdf_merge(const float* lengths, const float* weights, const float* u, int64_t u_row_stride,
                        const int64_t* rng_state, int rng_site, float* new_lengths, int64_t* inds, int32_t* flag,
                        int64_t R, int P, int n_new, int add_input_samples, void* stream);

/* Plain sample_pdf_python (renderers/utils.py:83-158) on explicit bin edges: bins [R,n_bins],
 * weights [R,n_bins-1] -> samples [R,n_samples] in draw order (not sorted), same rounding contract. */
int yn_sample_pdf(const float* bins, const float* weights, const float* u, int64_t u_row_stride, float* samples,
                  int64_t* inds, int32_t* flag, int64_t R, int n_bins, int n_samples, void* stream);

/* ------------------------------------------------------------------------------------------------
 * In-kernel draws of the training step.  rng_state: int64[4] in DEVICE memory = {seed, steps begun,
 * current step (snapshot), reserved}.  Philox4x32-10, key = seed (+ step high word), counter = (row, group of four
 * consecutive elements, draw site, step): uniforms carry 24 random bits like torch.rand, normals are Box-Muller pairs.
 * Reference draw sites: stratified jitter rand_like (ray_sampler.py:384), density noise randn_like
 * (multipass_emission_absorpsion_renderer.py:203-207), inverse-CDF uniforms torch.rand (renderers/utils.py:133-134).
 * An explicit draw pointer always takes precedence (parity tests replay the reference's draws through it).
 * ---------------------------------------------------------------------------------------------- */
/* One thread: rng_state[2] = rng_state[1]++ (the snapshot every kernel of the step keys its draws with) and
 * adam_state[0] += 1 (the float step counter yn_adam_step_dev reads).  Either pointer may be NULL. */
int yn_step_begin(int64_t* rng_state, float* adam_state, void* stream);

/* Training rays in ONE launch: the unmasked pixel pick of yn_sample_pixels (keyed by rng_state), the pinhole rays of
 * yn_ray_bundle and the stratified depths with the jitter drawn in the kernel (site rng_site) when stratified != 0.
 *   out: idx_out int64 [B,n] (may be NULL), xy_out [B,n,2], origins / directions [B,n,3], lengths [B,n,P] */
int yn_train_rays(const int64_t* rng_state, int rng_site, const float* poses, int64_t pose_batch_stride,
                  int64_t pose_row_stride, const float* focal, const float* depths, int stratified, int64_t* idx_out,
                  float* xy_out, float* origins, float* directions, float* lengths, int64_t B, int64_t n, int P,
                  int width, int height, void* stream);

/* The draws the kernels generate for (rng_site, current step), written out: out [R,P]; kind 0 = U[0,1), 1 = N(0,1).
 * Feeding them back through the explicit draw pointers reproduces the in-kernel path bit for bit (tests). */
int yn_rng_fill(const int64_t* rng_state, int rng_site, int kind, float* out, int64_t R, int P, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Per-image rgb losses: ground-truth gather at the sampled pixels (sample_grid, pipelines/utils.py:272-296:
 * flat index = (x + width * y).long()) fused with _rgb_metrics (137-158): mse[b] = mean over (rays x channels) of
 * (pred - gt)^2 and huber[b] = (sqrt(max(1 + mse / 0.03^2, 0) + 1e-4) - 1) * 0.03 (189-203).
 *   pred [B,n,C], image [B,height,width,C], xy [B,n,2] -> mse [B], huber [B]; deterministic reduction order.
 * Backward: d_pred [B,n,C] for incoming g_mse [B], g_huber [B] (either may be NULL).
 * ---------------------------------------------------------------------------------------------- */
/* scratch: yn_rgb_loss_scratch_bytes(B) bytes of device memory, zeroed ONCE by the caller (partial sums + a per-image
 * block counter that the kernel resets itself); it may be reused by later calls on the same stream. */
int64_t yn_rgb_loss_scratch_bytes(int64_t B);
int yn_rgb_loss_fwd(const float* pred, const float* image, const float* xy, float* mse, float* huber, void* scratch,
                    int64_t B, int64_t n, int C, int width, int height, void* stream);
int yn_rgb_loss_bwd(const float* pred, const float* image, const float* xy, const float* mse, const float* g_mse,
                    const float* g_huber, float* d_pred, int64_t B, int64_t n, int C, int width, int height,
                    void* stream);

/* scatter_rays_to_image (pipelines/utils.py:299-323) for up to three tensors at once: src[k] [B,n,channels[k]] is
 * written to the CALLER-ZEROED canvas dst[k] [B,height,width,channels[k]] at flat pixel (x + width * y).long().
 * src / dst / channels are HOST arrays of n_tensors entries. */
int yn_scatter_rays(const float* const* src, float* const* dst, const int* channels, int n_tensors, const float* xy,
                    int64_t B, int64_t n, int width, int height, void* stream);

/* ------------------------------------------------------------------------------------------------
 * torch.optim.Adam step (scripts/run.py:159; weight_decay 0, amsgrad off) on flat buffers; grad_scale
 * multiplies the gradient first (1/world_size after a sum all-reduce, or 1/loss_scale).  beta1 / beta2 are
 * doubles: the bias corrections 1 - beta^step are evaluated in double precision like torch does.
 * ---------------------------------------------------------------------------------------------- */
int yn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                 double beta1, double beta2, float eps, int32_t step, float grad_scale, void* stream);

/* Same update with the step counter and learning rate read from DEVICE memory (state[0] = step as float >= 1,
 * state[1] = lr): the launch can be part of a captured CUDA graph. */
int yn_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                     const float* state, double beta1, double beta2, float eps, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* YANERF_B200_H_ */
