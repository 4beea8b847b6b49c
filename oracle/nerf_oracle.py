"""CPU oracle for the yet-another-nerf per-ray render-and-train hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
module: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the checker
or as the timed CPU baseline, never as the thing shipped.  (Every function follows the
device of its inputs, so ``bench.py`` can also time the same restatement as
"torch eager on the GPU": the stronger baseline of SURVEY 8(d) / BASELINE.md 3(c).)

It is a from-scratch functional restatement (flat ``[R, P]`` tensors, explicit
random draws passed in as arguments, no nn.Module state) of the reference's
torch algorithm.  Each function cites the reference lines it follows
(paths relative to ``/root/reference``).  The arithmetic primitives are torch's
CPU fp32 ops, which is also what the reference runs with ``--device cpu``.

Parity is PINNED: ``tests/golden/make_golden.py`` executes the unmodified
reference in the build container (with import shims for addict/yapf/imageio/
omegaconf/torch._six) on seeded inputs and commits its outputs to
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` requires this oracle to
reproduce them bit-for-bit.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------- #
# configuration records (plain data; mirror the YAML keys of configs/nerf/*.yml)
# --------------------------------------------------------------------------- #
@dataclass
class MLPSpec:
    """Architecture of one NeRFMLP (`models/nerf_mlp.py:14-30`)."""

    n_layers: int = 8
    input_skips: Tuple[int, ...] = (5,)
    n_harmonic_functions_xyz: int = 10
    n_hidden_neurons_xyz: int = 256
    n_harmonic_functions_dir: int = 4
    n_hidden_neurons_dir: int = 128
    color_dim: int = 3
    latent_dim: int = 0  # width of the per-image global code appended to the xyz embedding (nerf_mlp.py:24, 86)

    @property
    def embed_xyz(self) -> int:  # models/utils.py:122
        return 3 * (2 * self.n_harmonic_functions_xyz + 1)

    @property
    def embed_dir(self) -> int:
        return 3 * (2 * self.n_harmonic_functions_dir + 1)

    def layer_dims(self) -> List[Tuple[int, int]]:
        """(in, out) of every trunk layer (`nerf_mlp.py:244-254`)."""
        # NB: `_construct_xyz_encoder` (nerf_mlp.py:88-95) does not forward
        # hidden_dim, so inner layers are always 256 wide (MLPWithInputSkips
        # default, nerf_mlp.py:225); only the LAST layer emits n_hidden_neurons_xyz.
        hidden = 256
        dims = []
        for li in range(self.n_layers):
            emb = self.embed_xyz + self.latent_dim  # nerf_mlp.py:85-86: embedding + global code
            din = hidden if li > 0 else emb
            if li > 0 and li in self.input_skips:
                din = hidden + emb
            dims.append((din, hidden if li + 1 < self.n_layers else self.n_hidden_neurons_xyz))
        return dims

    def param_shapes(self) -> "Dict[str, Tuple[int, ...]]":
        """State-dict keys (relative to the NeRFMLP module) and shapes, in
        registration order (`nerf_mlp.py:61-83`, SURVEY §5 checkpoint)."""
        out: Dict[str, Tuple[int, ...]] = {}
        for li, (din, dout) in enumerate(self.layer_dims()):
            out[f"xyz_encoder.mlp.{li}.0.weight"] = (dout, din)
            out[f"xyz_encoder.mlp.{li}.0.bias"] = (dout,)
        h, hd = self.n_hidden_neurons_xyz, self.n_hidden_neurons_dir
        out["intermediate_linear.weight"] = (h, h)
        out["intermediate_linear.bias"] = (h,)
        out["density_layer.weight"] = (1, h)
        out["density_layer.bias"] = (1,)
        out["color_layer.0.weight"] = (hd, h + self.embed_dir)
        out["color_layer.0.bias"] = (hd,)
        out["color_layer.2.weight"] = (self.color_dim, hd)
        out["color_layer.2.bias"] = (self.color_dim,)
        return out


@dataclass
class RaymarcherSpec:
    """`EmissionAbsorptionRaymarcher` ctor as built by the multipass renderer
    (`multipass_emission_absorpsion_renderer.py:44-53,120-152`)."""

    background_opacity: float = 1e10
    background_density_bias: float = 0.0
    blend_output: bool = False
    hard_background: bool = False
    bg_color: Tuple[float, ...] = (0.0,)


# --------------------------------------------------------------------------- #
# kernel family 1: ray sampler + harmonic encoding
# --------------------------------------------------------------------------- #
def xy_grid(height: int, width: int) -> Tensor:
    """Float pixel coordinates stacked (x, y), shape [H, W, 2]
    (`ray_samplers/utils.py:12-24`)."""
    ys = torch.linspace(0, height - 1, height, dtype=torch.float32)
    xs = torch.linspace(0, width - 1, width, dtype=torch.float32)
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack((gx, gy), dim=-1)


def rays_from_xy(poses: Tensor, focal: Tensor, xy: Tensor, width: int, height: int) -> Tuple[Tensor, Tensor]:
    """origins/directions for pixel coordinates `xy` [B, n, 2]
    (`ray_sampler.py:297-312`).  Directions are NOT normalised."""
    B, n, _ = xy.shape
    pose = poses[:, :3, :4]
    origins = pose[:, None, :, 3].expand(B, n, 3)
    f = focal.reshape(B, 1)
    cam = torch.stack(((xy[..., 0] - width * 0.5) / f, (xy[..., 1] - height * 0.5) / f, torch.ones(B, n, device=xy.device)), dim=-1)
    directions = torch.sum(pose[:, None, :, :3] * cam[:, :, None, :], dim=-1)
    return origins, directions


def depth_linspace(min_depth: float, max_depth: float, n_pts: int) -> Tensor:
    """`ray_sampler.py:285-291`."""
    return torch.linspace(min_depth, max_depth, n_pts, dtype=torch.float32)


def stratified_jitter(z: Tensor, u: Tensor) -> Tensor:
    """`_jiggle_within_stratas` with the uniform draw `u` made explicit
    (`ray_sampler.py:361-386`)."""
    mids = 0.5 * (z[..., 1:] + z[..., :-1])
    upper = torch.cat((mids, z[..., -1:]), dim=-1)
    lower = torch.cat((z[..., :1], mids), dim=-1)
    return lower + (upper - lower) * u


def harmonic_embedding(x: Tensor, n_freq: int) -> Tensor:
    """sin | cos | x with frequencies 2^k, channel order x·f0..x·f{L-1}, y·…
    (`models/utils.py:74-78,90-103`)."""
    freqs = 2.0 ** torch.arange(n_freq, dtype=torch.float32, device=x.device)
    e = (x[..., None] * freqs).reshape(*x.shape[:-1], -1)
    return torch.cat((e.sin(), e.cos(), x), dim=-1)


# --------------------------------------------------------------------------- #
# kernel family 2: the NeRF MLP
# --------------------------------------------------------------------------- #
def mlp_forward(
    params: Dict[str, Tensor], spec: MLPSpec, origins: Tensor, directions: Tensor, lengths: Tensor,
    global_codes: Optional[Tensor] = None,
) -> Tuple[Tensor, Tensor]:
    """`NeRFMLP.forward` (`nerf_mlp.py:117-177`) on flat rays.

    origins/directions [R, 3], lengths [R, P] -> raw density [R, P], rgb [R, P, 3].
    global_codes [R, latent_dim]: the ray's image code (`broadcast_global_code`, nerf_mlp.py:324-335: appended to the
    xyz embedding of every point of the image), required iff spec.latent_dim > 0 (`_check_input`, 179-183).
    """
    if (global_codes is None) != (spec.latent_dim == 0) or (global_codes is not None and global_codes.shape[-1] != spec.latent_dim):
        raise ValueError("The shape of global codes is imcompible with the input dim of the network.")
    pts = origins[:, None, :] + lengths[:, :, None] * directions[:, None, :]  # models/utils.py:244
    emb = harmonic_embedding(pts, spec.n_harmonic_functions_xyz)
    if global_codes is not None:
        emb = torch.cat((emb, global_codes[:, None, :].expand(-1, emb.shape[1], -1)), dim=-1)
    y = emb
    for li in range(spec.n_layers):  # nerf_mlp.py:281-288
        if li in spec.input_skips:
            y = torch.cat((y, emb), dim=-1)
        y = torch.relu(F.linear(y, params[f"xyz_encoder.mlp.{li}.0.weight"], params[f"xyz_encoder.mlp.{li}.0.bias"]))
    raw_density = F.linear(y, params["density_layer.weight"], params["density_layer.bias"])[..., 0]
    # nerf_mlp.py:97-115 and LinearWithRepeat models/utils.py:207-211
    demb = harmonic_embedding(F.normalize(directions, dim=-1), spec.n_harmonic_functions_dir)
    inter = F.linear(y, params["intermediate_linear.weight"], params["intermediate_linear.bias"])
    h = spec.n_hidden_neurons_xyz
    wc = params["color_layer.0.weight"]
    hid = F.linear(inter, wc[:, :h], params["color_layer.0.bias"]) + F.linear(demb, wc[:, h:], None)[:, None, :]
    hid = torch.relu(hid)
    rgb = torch.sigmoid(F.linear(hid, params["color_layer.2.weight"], params["color_layer.2.bias"]))
    return raw_density, rgb


# --------------------------------------------------------------------------- #
# kernel family 3: emission-absorption compositing
# --------------------------------------------------------------------------- #
def raymarch(
    raw_density: Tensor,
    rgb: Tensor,
    lengths: Tensor,
    directions: Tensor,
    spec: RaymarcherSpec,
    noise: Optional[Tensor] = None,
    noise_std: float = 0.0,
    bg_color: Optional[Tensor] = None,
) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """`EmissionAbsorptionRaymarcher.forward`
    (`multipass_emission_absorpsion_renderer.py:194-239`), exponential capping,
    product weights, surface_thickness 1.

    raw_density [R,P], rgb [R,P,C], lengths [R,P], directions [R,3],
    noise [R,P] standard normal draw (used when noise_std > 0), bg_color [R,C']
    -> features [R,C], depths [R,1], opacities [R,1], weights [R,P].
    """
    deltas = torch.cat(
        (lengths[:, 1:] - lengths[:, :-1], spec.background_opacity * torch.ones_like(lengths[:, :1])), dim=-1
    )
    deltas = deltas * directions[:, None, :].norm(p=2, dim=-1)
    dens = raw_density
    if noise_std > 0.0:
        dens = dens + noise * noise_std
    dens = torch.relu(dens) + spec.background_density_bias
    x = deltas * dens
    alpha = 1.0 - torch.exp(-x)
    opac_run = 1.0 - torch.exp(-torch.cumsum(x, dim=-1))
    opacities = opac_run[:, -1:]
    trans = (1.0 - opac_run).roll(1, dims=-1)
    trans[:, :1] = 1.0
    weights = alpha * trans
    depths = (weights * lengths)[..., None].sum(dim=-2)
    if bg_color is None:
        bg = torch.tensor(spec.bg_color, dtype=torch.float32, device=rgb.device).view(1, -1).expand(rgb.shape[0], -1)
    else:
        bg = bg_color
    if not spec.hard_background:
        feats = (weights[..., None] * rgb).sum(dim=-2)
        if bg.shape[-1] not in (1, feats.shape[-1]):
            raise ValueError("Wrong number of background color channels")
        a = opacities if spec.blend_output else 1
        feats = a * feats + (1 - opacities) * bg
    else:
        rgb = torch.cat([rgb[:, :-1, :], bg[:, None, :]], dim=-2)
        feats = (weights[..., None] * rgb).sum(dim=-2)
    return feats, depths, opacities, weights


# --------------------------------------------------------------------------- #
# kernel family 4: inverse-CDF resampling + merge
# --------------------------------------------------------------------------- #
def sample_pdf(
    bins: Tensor, weights: Tensor, n_samples: int, u: Optional[Tensor] = None, eps: float = 1e-5
) -> Tuple[Tensor, Tensor]:
    """`sample_pdf_python` (`renderers/utils.py:121-158`).

    `u` None means the deterministic `linspace(0,1,n)` branch (130-132); otherwise
    `u` [R, n] stands for the `torch.rand` draw (133-134).  Returns the samples
    and the searchsorted indices (line 137) so index parity can be asserted.
    """
    w = weights + eps
    if w.min() <= 0:
        raise ValueError("Negative weights provided.")
    pdf = w / w.sum(dim=-1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    if u is None:
        u = torch.linspace(0.0, 1.0, n_samples, dtype=cdf.dtype).to(cdf.device)  # evaluated on the CPU like the reference
        u = u.expand(list(cdf.shape[:-1]) + [n_samples]).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = (inds - 1).clamp(0)
    above = inds.clamp(max=cdf.shape[-1] - 1)
    cdf_b, cdf_a = torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)
    bin_b, bin_a = torch.gather(bins, -1, below), torch.gather(bins, -1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < eps, torch.ones_like(denom), denom)
    t = (u - cdf_b) / denom
    return bin_b + t * (bin_a - bin_b), inds


def refine_lengths(
    lengths: Tensor, weights: Tensor, n_fine: int, u: Optional[Tensor], add_input_samples: bool = True
) -> Tuple[Tensor, Tensor]:
    """`RayPointRefiner.forward` (`renderers/utils.py:48-69`) on flat rays."""
    with torch.no_grad():  # renderers/utils.py:50 -- no gradient reaches the coarse weights
        mid = torch.lerp(lengths[:, 1:], lengths[:, :-1], 0.5)
        z_new, inds = sample_pdf(mid, weights[:, 1:-1], n_fine, u)
    z = torch.cat((lengths, z_new), dim=-1) if add_input_samples else z_new
    z, _ = torch.sort(z, dim=-1)
    return z, inds


# --------------------------------------------------------------------------- #
# per-sample losses (`pipelines/utils.py:137-203`)
# --------------------------------------------------------------------------- #
def rgb_losses(pred: Tensor, gt: Tensor) -> Tuple[Tensor, Tensor]:
    """pred, gt [B, n, 3] -> (mse [B], huber [B])."""
    B = pred.shape[0]
    mse = ((pred.reshape(B, -1) - gt.reshape(B, -1)) ** 2).mean(dim=-1)
    s = 0.03
    hub = ((torch.clamp(1 + mse / (s * s), 0.0) + 1e-4).sqrt() - 1) * s
    return mse, hub


def gather_pixels(image: Tensor, xy: Tensor) -> Tensor:
    """`sample_grid` (`pipelines/utils.py:272-296`): image [B,H,W,C], xy [B,n,2]."""
    B, H, W, C = image.shape
    idx = (xy[..., 0] + W * xy[..., 1]).long()
    return torch.gather(image.reshape(B, H * W, C), 1, idx[..., None].expand(-1, -1, C))


def chunk_plan(n_rays: int, n_pts: int, chunk_size: int) -> Tuple[int, int]:
    """(n_chunks, rays per chunk) of the chunkify contract (`nerf_pipeline.py:349-352`)."""
    n_chunks = -(-n_rays * max(n_pts, 1) // chunk_size)
    return n_chunks, -(-n_rays // n_chunks)


# --------------------------------------------------------------------------- #
# the composed hot path
# --------------------------------------------------------------------------- #
@dataclass
class PipelineSpec:
    """The keys of `configs/nerf/lego.yml:45-94` that reach the hot path."""

    mlp: MLPSpec = field(default_factory=MLPSpec)
    march: RaymarcherSpec = field(default_factory=lambda: RaymarcherSpec(background_density_bias=1e-6, bg_color=(0.0, 0.0, 0.0)))
    image_height: int = 800
    image_width: int = 800
    min_depth: float = 2.0
    max_depth: float = 6.0
    n_pts_coarse: int = 64
    n_pts_fine: int = 128
    density_noise_std_train: float = 0.2
    append_coarse_samples_to_fine: bool = True
    chunk_size_grid: int = 131072
    loss_weights: Dict[str, float] = field(
        default_factory=lambda: {"loss_prev_stage_rgb_mse": 1.0, "loss_rgb_mse": 1.0}
    )


def render_rays(
    nets: Sequence[Dict[str, Tensor]],
    spec: PipelineSpec,
    origins: Tensor,
    directions: Tensor,
    lengths: Tensor,
    training: bool,
    draws: Optional[Dict[str, Tensor]] = None,
    bg_color: Optional[Tensor] = None,
) -> List[Dict[str, Tensor]]:
    """`MultipassEmissionAbsorpsionRenderer._run_raymarcher`
    (`multipass_emission_absorpsion_renderer.py:84-117`) for flat rays.

    Returns one dict per pass (coarse first) with features/depths/opacities/
    weights/lengths (+ `inds` from the refiner on all but the last pass).
    `draws` keys (training only): noise0, u_pdf, noise1.
    """
    std = spec.density_noise_std_train if training else 0.0
    outs: List[Dict[str, Tensor]] = []
    z = lengths
    for k, net in enumerate(nets):
        dens, rgb = mlp_forward(net, spec.mlp, origins, directions, z)
        noise = draws[f"noise{k}"] if (training and std > 0.0) else None
        f, d, o, w = raymarch(dens, rgb, z, directions, spec.march, noise, std, bg_color)
        rec = dict(features=f, depths=d, opacities=o, weights=w, lengths=z, raw_density=dens, rgb=rgb)
        if k + 1 < len(nets):
            u = draws["u_pdf"] if training else None
            z, inds = refine_lengths(z, w, spec.n_pts_fine, u, spec.append_coarse_samples_to_fine)
            rec["inds"] = inds
        outs.append(rec)
    return outs


def train_forward(
    nets: Sequence[Dict[str, Tensor]],
    spec: PipelineSpec,
    poses: Tensor,
    focal: Tensor,
    image_rgb: Tensor,
    draws: Dict[str, Tensor],
) -> Dict[str, Tensor]:
    """`NeRFPipeline.forward(evaluation_mode=TRAINING)` (`nerf_pipeline.py:138-215`).

    `draws`: pix [B,n] int64 (the multinomial result), u_strat [B,n,Pc],
    noise0 [B*n,Pc], u_pdf [B*n,Nf], noise1 [B*n,Pc+Nf] (SURVEY §8(d) call order).
    """
    B = poses.shape[0]
    H, W = spec.image_height, spec.image_width
    grid = xy_grid(H, W).to(poses.device).reshape(1, H * W, 2).expand(B, -1, -1)
    xy = torch.gather(grid, 1, draws["pix"][..., None].expand(-1, -1, 2))
    n = xy.shape[1]
    o, d = rays_from_xy(poses, focal, xy, W, H)
    z = depth_linspace(spec.min_depth, spec.max_depth, spec.n_pts_coarse).to(poses.device)[None, None].expand(B, n, -1)
    z = stratified_jitter(z, draws["u_strat"])
    passes = render_rays(nets, spec, o.reshape(-1, 3), d.reshape(-1, 3), z.reshape(B * n, -1), True, draws)
    gt = gather_pixels(image_rgb, xy)
    out: Dict[str, Tensor] = {"xys": xy}
    prefixes = ["loss_prev_stage_", "loss_"] if len(passes) == 2 else ["loss_"]
    for pref, rec in zip(prefixes[-len(passes):], passes):
        mse, hub = rgb_losses(rec["features"].reshape(B, n, -1), gt)
        out[pref + "rgb_mse"], out[pref + "rgb_huber"] = mse, hub
    out["objective"] = sum(out[k] * float(w) for k, w in spec.loss_weights.items() if k in out and w != 0.0)
    out["passes"] = passes
    return out


def render_image(
    nets: Sequence[Dict[str, Tensor]],
    spec: PipelineSpec,
    poses: Tensor,
    focal: Tensor,
    ray_slice: Optional[Tuple[int, int]] = None,
) -> Dict[str, Tensor]:
    """`NeRFPipeline.forward(evaluation_mode=EVALUATION)` with the chunk loop of
    `_chunk_generator` (`nerf_pipeline.py:333-377`); `ray_slice` restricts to rays
    [start, end) of the flattened grid (used for bounded CPU timing)."""
    B = poses.shape[0]
    H, W = spec.image_height, spec.image_width
    xy = xy_grid(H, W).to(poses.device).reshape(1, H * W, 2).expand(B, -1, -1)
    if ray_slice is not None:
        xy = xy[:, ray_slice[0]:ray_slice[1]]
    n = xy.shape[1]
    o, d = rays_from_xy(poses, focal, xy, W, H)
    z = depth_linspace(spec.min_depth, spec.max_depth, spec.n_pts_coarse).to(poses.device)[None, None].expand(B, n, -1)
    _, per = chunk_plan(n, spec.n_pts_coarse, spec.chunk_size_grid) if spec.chunk_size_grid > 0 else (1, n)
    acc: Dict[str, List[Tensor]] = {"features": [], "depths": [], "opacities": [], "coarse_features": []}
    for s in range(0, n, per):
        e = min(s + per, n)
        passes = render_rays(
            nets, spec, o[:, s:e].reshape(-1, 3), d[:, s:e].reshape(-1, 3), z[:, s:e].reshape(B * (e - s), -1), False
        )
        for k in ("features", "depths", "opacities"):
            acc[k].append(passes[-1][k].reshape(B, e - s, -1))
        acc["coarse_features"].append(passes[0]["features"].reshape(B, e - s, -1))
    return {k: torch.cat(v, dim=1) for k, v in acc.items()}


def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float,
              beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8) -> None:
    """torch.optim.Adam single-tensor update (the optimizer `scripts/run.py:159`
    builds; weight_decay 0, amsgrad False).  In place on p, m, v."""
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))
