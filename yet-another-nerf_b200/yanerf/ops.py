"""Operator layer: torch-facing wrappers (and autograd Functions) over the C ABI.

Every function takes/returns CUDA fp32 tensors with rays flattened to `[R, ...]`.
Nothing here computes on the CPU or with torch ops: the arithmetic happens in
libyanerf_b200.so; torch only owns the memory and the stream.
"""
from __future__ import annotations

import contextlib
import ctypes
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import torch

from . import _native as N


class Profiler:
    """Optional per-kernel CUDA-event timing (bench.py / tools): events are recorded on the launching
    stream around each C-ABI call; `launches` counts device kernels launched by this package."""

    enabled = False
    records: list = []
    launches = 0
    KERNELS_PER_CALL = {"yn_mlp_pack_weights": 3, "yn_mlp_bwd": 6}  # bwd: gradient amax, dgrad, wgrad, inter, heads, direction

    @classmethod
    def reset(cls):
        cls.records, cls.launches = [], 0

    @classmethod
    def summary(cls):
        """name -> (calls, total ms); call after torch.cuda.synchronize()."""
        out = {}
        for name, s, e in cls.records:
            c, t = out.get(name, (0, 0.0))
            out[name] = (c + 1, t + s.elapsed_time(e))
        return out


STREAM = object()  # placeholder argument: the current stream of the launch device, resolved inside `_call`


def _call(name: str, *args, device=None) -> None:
    """Launch one C-ABI entry point on `device`'s current stream (the tensors' device, which need not be torch's
    current device).  Callers keep every converted operand in a local variable until this returns: the launch is
    stream-ordered, so a temporary may only go back to the caching allocator after the kernel has been queued."""
    fn = getattr(N.lib(), name)
    current = torch.cuda.current_device()
    index = current if (device is None or device.index is None) else device.index
    stream = N.stream_ptr(index)
    args = tuple(stream if a is STREAM else a for a in args)
    with (contextlib.nullcontext() if index == current else torch.cuda.device(index)):
        if Profiler.enabled:
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            rc = fn(*args)
            e.record()
            Profiler.records.append((name, s, e))
        else:
            rc = fn(*args)
    Profiler.launches += Profiler.KERNELS_PER_CALL.get(name, 1)
    N.check(rc)


# --------------------------------------------------------------------------- #
# ray sampler
# --------------------------------------------------------------------------- #
def ray_bundle(
    poses: torch.Tensor,
    focal: torch.Tensor,
    xy: Optional[torch.Tensor],
    depths: torch.Tensor,
    u: Optional[torch.Tensor],
    n_rays: int,
    width: int,
    height: int,
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """`_xy_to_ray_bundle` + `_jiggle_within_stratas` (ray_sampler.py:249-314, 361-386).

    poses [B,3,4] (any strides with unit column stride), focal [B], xy [B,n,2] or
    None for the full `height x width` grid, depths [P], u [B,n,P] or None.
    Returns origins [B,n,3], directions [B,n,3], lengths [B,n,P], xys [B,n,2].
    """
    B = poses.shape[0]
    P = depths.shape[0]
    dev = poses.device
    if poses.dtype != torch.float32 or poses.stride(2) != 1:
        poses = poses.float().contiguous()
    focal = N.f32c(focal.reshape(B))
    full = xy is None
    if not full:
        xy = N.f32c(xy)
    origins = torch.empty(B, n_rays, 3, device=dev)
    directions = torch.empty(B, n_rays, 3, device=dev)
    lengths = torch.empty(B, n_rays, P, device=dev)
    xys = torch.empty(B, n_rays, 2, device=dev) if full else xy
    depths = N.f32c(depths)
    u = None if u is None else N.f32c(u)
    _call("yn_ray_bundle",
          ctypes.c_void_p(poses.data_ptr()), poses.stride(0), poses.stride(1), N.ptr(focal), N.ptr(xy),
          N.ptr(depths), N.ptr(u), N.ptr(origins), N.ptr(directions),
          N.ptr(lengths), N.ptr(xys if full else None), B, n_rays, P, width, height, 1 if full else 0,
          STREAM, device=dev)
    return origins, directions, lengths, xys


def sample_pixels(seed: torch.Tensor, batch: int, n: int, width: int, height: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """n distinct uniformly random pixels per image (the unmasked `torch.multinomial` pick of the reference's
    ray sampler).  seed: int64[1] CUDA tensor.  Returns idx int64 [B,n] and xy float [B,n,2]."""
    idx = torch.empty(batch, n, dtype=torch.int64, device=seed.device)
    xy = torch.empty(batch, n, 2, device=seed.device)
    _call("yn_sample_pixels", N.ptr(seed, torch.int64), N.ptr(idx, torch.int64), N.ptr(xy), batch, n, width, height,
          STREAM, device=seed.device)
    return idx, xy


class DeviceRng:
    """Device-resident generator state of the in-kernel draws (`int64[4]`: seed, steps begun, current step, reserved;
    include/yanerf_b200.h "in-kernel draws").  `step_begin` snapshots the step; every kernel of the step keys its
    Philox counters with (row, element group, site, step), so nothing random round-trips through HBM and a captured
    CUDA graph advances by itself.  Sites: stratified jitter 0, density noise 16 + pass, inverse-CDF uniforms 32 + pass."""

    SITE_STRATIFIED, SITE_NOISE, SITE_PDF = 0, 16, 32

    def __init__(self, device, seed: Optional[int] = None) -> None:
        if seed is None:  # from torch's CPU generator: reproducible under torch.manual_seed, no device sync
            seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64))
        self.state = torch.tensor([seed, 0, 0, 0], dtype=torch.int64).to(device)

    def seed(self, seed: int, step: int = 0) -> None:
        self.state.copy_(torch.tensor([seed, step, max(step - 1, 0), 0], dtype=torch.int64))


def rng_fill(rng: DeviceRng, site: int, rows: int, cols: int, normal: bool) -> torch.Tensor:
    """The `[rows, cols]` draws the kernels generate for `site` in the current step (tests / debugging)."""
    out = torch.empty(rows, cols, device=rng.state.device)
    _call("yn_rng_fill", N.ptr(rng.state, torch.int64), site, int(bool(normal)), N.ptr(out), rows, cols, STREAM,
          device=out.device)
    return out


def step_begin(rng: Optional[DeviceRng], adam_state: Optional[torch.Tensor]) -> None:
    """One launch at the start of a training step: snapshot the step for the draws, bump Adam's device-side step."""
    dev = rng.state.device if rng is not None else adam_state.device
    _call("yn_step_begin", N.ptr(None if rng is None else rng.state, torch.int64), N.ptr(adam_state), STREAM, device=dev)


def train_rays(rng: DeviceRng, poses: torch.Tensor, focal: torch.Tensor, depths: torch.Tensor, stratified: bool,
               n_rays: int, width: int, height: int):
    """Unmasked training pixel pick + `_xy_to_ray_bundle` + `_jiggle_within_stratas` in one launch
    (ray_sampler.py:187-229, 249-314, 361-386), jitter drawn in the kernel.
    Returns idx int64 [B,n], xys [B,n,2], origins [B,n,3], directions [B,n,3], lengths [B,n,P]."""
    B, P, dev = poses.shape[0], depths.shape[0], poses.device
    if poses.dtype != torch.float32 or poses.stride(2) != 1:
        poses = poses.float().contiguous()
    focal = N.f32c(focal.reshape(B))
    depths = N.f32c(depths)
    idx = torch.empty(B, n_rays, dtype=torch.int64, device=dev)
    xys = torch.empty(B, n_rays, 2, device=dev)
    origins = torch.empty(B, n_rays, 3, device=dev)
    directions = torch.empty(B, n_rays, 3, device=dev)
    lengths = torch.empty(B, n_rays, P, device=dev)
    _call("yn_train_rays", N.ptr(rng.state, torch.int64), DeviceRng.SITE_STRATIFIED, ctypes.c_void_p(poses.data_ptr()),
          poses.stride(0), poses.stride(1), N.ptr(focal), N.ptr(depths), int(bool(stratified)), N.ptr(idx, torch.int64),
          N.ptr(xys), N.ptr(origins), N.ptr(directions), N.ptr(lengths), B, n_rays, P, width, height, STREAM, device=dev)
    return idx, xys, origins, directions, lengths


# --------------------------------------------------------------------------- #
# NeRF MLP
# --------------------------------------------------------------------------- #
@dataclass
class MlpPlan:
    """Device-side state of one NeRFMLP: architecture + packed tensor-core weights."""

    arch: N.MlpArch
    n_params: int
    wpack: torch.Tensor  # uint8
    aux: torch.Tensor  # fp32

    @staticmethod
    def create(arch: N.MlpArch, device) -> "MlpPlan":
        L = N.lib()
        n_params = L.yn_mlp_param_count(ctypes.byref(arch))
        if n_params < 0:
            raise NotImplementedError(L.yn_last_error_string().decode())
        wbytes = L.yn_mlp_wpack_bytes(ctypes.byref(arch))
        naux = L.yn_mlp_aux_floats(ctypes.byref(arch))
        return MlpPlan(
            arch=arch,
            n_params=int(n_params),
            wpack=torch.empty(int(wbytes), dtype=torch.uint8, device=device),
            aux=torch.empty(int(naux), dtype=torch.float32, device=device),
        )

    def pack(self, flat_params: torch.Tensor) -> None:
        assert flat_params.numel() == self.n_params
        _call("yn_mlp_pack_weights", ctypes.byref(self.arch), N.ptr(flat_params), N.ptr(self.wpack, torch.uint8),
              N.ptr(self.aux), STREAM, device=self.wpack.device)


def mlp_forward_raw(
    plan: MlpPlan, flat_params: torch.Tensor, origins: torch.Tensor, directions: torch.Tensor, lengths: torch.Tensor,
    stash: Optional[torch.Tensor] = None,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """origins/directions [R,3], lengths [R,P] -> raw density [R,P], rgb [R,P,C] (plan already packed)."""
    R, P = lengths.shape
    dev = lengths.device
    C = plan.arch.color_dim
    density = torch.empty(R, P, device=dev)
    rgb = torch.empty(R, P, C, device=dev)
    if R == 0:
        return density, rgb
    dirbias = torch.empty(R, 128, device=dev)
    L = N.lib()
    _call("yn_mlp_dirbias", ctypes.byref(plan.arch), N.ptr(flat_params), N.ptr(directions), N.ptr(dirbias), R, STREAM,
          device=dev)
    _call("yn_mlp_fwd",
          ctypes.byref(plan.arch), N.ptr(origins), N.ptr(directions), N.ptr(lengths), N.ptr(dirbias),
          N.ptr(plan.wpack, torch.uint8), N.ptr(plan.aux), N.ptr(density), N.ptr(rgb),
          N.ptr(stash, torch.uint8), R, P, STREAM, device=dev)
    return density, rgb


class MlpFunction(torch.autograd.Function):
    """autograd node of NeRFMLP.forward: inputs (origins, directions, lengths) carry no gradient
    (the reference never differentiates w.r.t. the rays); the flat parameter vector does."""

    @staticmethod
    def forward(ctx, flat_params, origins, directions, lengths, plan: MlpPlan, need_grad: bool, grad_out=None):
        # NB: grad mode is off inside Function.forward, so the caller decides whether to keep the stash.
        # grad_out: optional flat fp32 buffer the backward kernels accumulate into directly (FusedTrainer's flat
        # gradient); the autograd gradient of flat_params is then None.
        R, P = lengths.shape
        ctx.grad_out = grad_out
        stash = None
        if need_grad and R > 0:
            nbytes = N.lib().yn_mlp_stash_bytes(ctypes.byref(plan.arch), R * P)
            stash = torch.empty(int(nbytes), dtype=torch.uint8, device=lengths.device)
        density, rgb = mlp_forward_raw(plan, flat_params, origins, directions, lengths, stash)
        ctx.plan = plan
        ctx.stash = stash
        ctx.save_for_backward(flat_params, directions, rgb)
        ctx.shape = (R, P)
        return density, rgb

    @staticmethod
    def backward(ctx, d_density, d_rgb):
        flat_params, directions, rgb = ctx.saved_tensors
        plan: MlpPlan = ctx.plan
        R, P = ctx.shape
        direct = ctx.grad_out is not None
        grads = ctx.grad_out if direct else torch.zeros_like(flat_params)
        if R > 0:
            L = N.lib()
            d_density = N.f32c(d_density) if d_density is not None else torch.zeros(R, P, device=rgb.device)
            d_rgb = N.f32c(d_rgb) if d_rgb is not None else torch.zeros_like(rgb)
            wbytes = L.yn_mlp_bwd_workspace_bytes(ctypes.byref(plan.arch), R * P)
            work = torch.empty(int(wbytes), dtype=torch.uint8, device=rgb.device)
            _call("yn_mlp_bwd",
                  ctypes.byref(plan.arch), N.ptr(directions), N.ptr(rgb), N.ptr(d_density), N.ptr(d_rgb),
                  N.ptr(flat_params), N.ptr(plan.wpack, torch.uint8), N.ptr(plan.aux), N.ptr(ctx.stash, torch.uint8),
                  N.ptr(work, torch.uint8), N.ptr(grads), R, P, STREAM, device=rgb.device)
        ctx.stash = None
        return (None if direct else grads), None, None, None, None, None, None


# --------------------------------------------------------------------------- #
# emission-absorption compositing
# --------------------------------------------------------------------------- #
def march_cfg(background_opacity: float, background_density_bias: float, density_noise_std: float, blend_output: bool,
              hard_background: bool, bg_const: Sequence[float]) -> N.MarchCfg:
    cfg = N.MarchCfg()
    cfg.background_opacity = background_opacity
    cfg.background_density_bias = background_density_bias
    cfg.density_noise_std = density_noise_std
    cfg.blend_output = int(blend_output)
    cfg.hard_background = int(hard_background)
    cfg.bg_channels = len(bg_const)
    for i, v in enumerate(bg_const):
        cfg.bg_const[i] = float(v)
    return cfg


class CompositeFunction(torch.autograd.Function):
    """EmissionAbsorptionRaymarcher.forward with its analytic backward."""

    @staticmethod
    def forward(ctx, raw_density, rgb, lengths, directions, noise, bg, cfg: N.MarchCfg, rng=None, site: int = 0):
        R, P = lengths.shape
        C = rgb.shape[-1]
        dev = lengths.device
        raw_density, rgb, lengths, directions = (N.f32c(t) for t in (raw_density, rgb, lengths, directions))
        noise = None if noise is None else N.f32c(noise)
        bg = None if bg is None else N.f32c(bg)
        if bg is not None:
            cfg = _copy_cfg(cfg)
            cfg.bg_channels = bg.shape[-1]
        features = torch.empty(R, C, device=dev)
        depths = torch.empty(R, 1, device=dev)
        opacities = torch.empty(R, 1, device=dev)
        weights = torch.empty(R, P, device=dev)
        rng_state = None if (rng is None or noise is not None) else rng.state
        _call("yn_composite_fwd",
              ctypes.byref(cfg), N.ptr(raw_density), N.ptr(rgb), N.ptr(lengths), N.ptr(directions), N.ptr(noise),
              N.ptr(rng_state, torch.int64), site,
              N.ptr(bg), N.ptr(features), N.ptr(depths), N.ptr(opacities), N.ptr(weights), R, P, C, STREAM, device=dev)
        ctx.cfg = cfg
        ctx.rng = (rng_state, site)
        ctx.has = (noise is not None, bg is not None)
        saved = [raw_density, rgb, lengths, directions] + ([noise] if noise is not None else []) + ([bg] if bg is not None else [])
        ctx.save_for_backward(*saved)
        return features, depths, opacities, weights

    @staticmethod
    def backward(ctx, d_features, d_depths, d_opacities, d_weights):
        saved = list(ctx.saved_tensors)
        raw_density, rgb, lengths, directions = saved[:4]
        rest = saved[4:]
        noise = rest.pop(0) if ctx.has[0] else None
        bg = rest.pop(0) if ctx.has[1] else None
        R, P = lengths.shape
        C = rgb.shape[-1]
        if d_features is None:
            d_features = torch.zeros(R, C, device=rgb.device)
        d_sigma = torch.empty_like(raw_density)
        d_rgb = torch.empty_like(rgb)
        # converted gradients stay referenced until the launch is queued (see `_call`)
        d_features, d_depths, d_opacities, d_weights = (None if t is None else N.f32c(t)
                                                        for t in (d_features, d_depths, d_opacities, d_weights))
        _call("yn_composite_bwd",
              ctypes.byref(ctx.cfg), N.ptr(raw_density), N.ptr(rgb), N.ptr(lengths), N.ptr(directions), N.ptr(noise),
              N.ptr(ctx.rng[0], torch.int64), ctx.rng[1],
              N.ptr(bg), N.ptr(d_features), N.ptr(d_depths), N.ptr(d_opacities),
              N.ptr(d_weights), N.ptr(d_sigma), N.ptr(d_rgb), R, P, C, STREAM, device=rgb.device)
        return d_sigma, d_rgb, None, None, None, None, None, None, None


def _copy_cfg(cfg: N.MarchCfg) -> N.MarchCfg:
    out = N.MarchCfg()
    ctypes.memmove(ctypes.byref(out), ctypes.byref(cfg), ctypes.sizeof(N.MarchCfg))
    return out


def composite(raw_density, rgb, lengths, directions, cfg: N.MarchCfg, noise=None, bg=None, rng: Optional[DeviceRng] = None,
              site: int = 0):
    """raw_density [R,P], rgb [R,P,C], lengths [R,P], directions [R,3] -> features [R,C], depths [R,1],
    opacities [R,1], weights [R,P].  Density noise (cfg.density_noise_std > 0): explicit `noise` [R,P], else drawn in the
    kernel from `rng` (same values in the backward kernel as long as the step has not advanced)."""
    if raw_density.shape[0] == 0:
        R, P = lengths.shape
        z = lengths.new_zeros
        return z(R, rgb.shape[-1]), z(R, 1), z(R, 1), z(R, P)
    return CompositeFunction.apply(raw_density, rgb, lengths, directions, noise, bg, cfg, rng, site)


# --------------------------------------------------------------------------- #
# inverse-CDF resampling + merge
# --------------------------------------------------------------------------- #
_LINSPACE_CACHE = {}


def det_draws(n: int, device) -> torch.Tensor:
    """`torch.linspace(0, 1, n)` exactly as the reference's --device cpu path computes it
    (renderers/utils.py:130-132); evaluated on the CPU once per (n, device) and cached."""
    key = (n, str(device))
    if key not in _LINSPACE_CACHE:
        _LINSPACE_CACHE[key] = torch.linspace(0.0, 1.0, n, dtype=torch.float32).to(device)
    return _LINSPACE_CACHE[key]


def sample_pdf_merge(lengths: torch.Tensor, weights: torch.Tensor, n_new: int, u: Optional[torch.Tensor],
                     add_input_samples: bool = True, want_inds: bool = False, flag: Optional[torch.Tensor] = None,
                     rng: Optional[DeviceRng] = None, site: int = 0):
    """lengths [R,P], weights [R,P] -> sorted new lengths [R, n_new (+P)], inds [R,n_new] int64 or None,
    flag int32[1] (1 if any weight + 1e-5 <= 0).  u: [R,n_new] draws; None = the deterministic linspace row, or, with
    `rng`, per-ray uniforms drawn in the kernel."""
    R, P = lengths.shape
    dev = lengths.device
    out = torch.empty(R, n_new + (P if add_input_samples else 0), device=dev)
    inds = torch.empty(R, n_new, dtype=torch.int64, device=dev) if want_inds else None
    if flag is None:
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
    if R == 0:
        return out, inds, flag
    rng_state = None
    if u is None and rng is not None:
        rng_state, stride = rng.state, 0
    elif u is None:
        u, stride = det_draws(n_new, dev), 0
    else:
        u = N.f32c(u)
        assert u.shape == (R, n_new), (u.shape, (R, n_new))
        stride = n_new
    lengths, weights = N.f32c(lengths), N.f32c(weights)
    _call("yn_sample_pdf_merge",
          N.ptr(lengths), N.ptr(weights), N.ptr(u), stride, N.ptr(rng_state, torch.int64), site, N.ptr(out),
          N.ptr(inds, torch.int64), N.ptr(flag, torch.int32), R, P, n_new, int(add_input_samples), STREAM, device=dev)
    return out, inds, flag


def sample_pdf(bins: torch.Tensor, weights: torch.Tensor, n_samples: int, u: Optional[torch.Tensor],
               want_inds: bool = False):
    """`sample_pdf_python` on explicit bins [R,nb] / weights [R,nb-1] -> samples [R,n] in draw order."""
    R, nb = bins.shape
    dev = bins.device
    out = torch.empty(R, n_samples, device=dev)
    inds = torch.empty(R, n_samples, dtype=torch.int64, device=dev) if want_inds else None
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    if R == 0:
        return out, inds, flag
    if u is None:
        u, stride = det_draws(n_samples, dev), 0
    else:
        u, stride = N.f32c(u), n_samples
    bins, weights = N.f32c(bins), N.f32c(weights)
    _call("yn_sample_pdf",
          N.ptr(bins), N.ptr(weights), N.ptr(u), stride, N.ptr(out), N.ptr(inds, torch.int64),
          N.ptr(flag, torch.int32), R, nb, n_samples, STREAM, device=dev)
    return out, inds, flag


# --------------------------------------------------------------------------- #
# per-image rgb losses (ground-truth gather + squared-error mean)
# --------------------------------------------------------------------------- #
_LOSS_SCRATCH = {}


def _loss_scratch(batch: int, device) -> torch.Tensor:
    """Zeroed-once device scratch of `yn_rgb_loss_fwd` (partial sums + block counters the kernel resets itself); one
    per (device, batch size): launches on a stream are ordered, so consecutive calls may share it."""
    key = (str(device), batch)
    if key not in _LOSS_SCRATCH:
        _LOSS_SCRATCH[key] = torch.zeros(int(N.lib().yn_rgb_loss_scratch_bytes(batch)), dtype=torch.uint8, device=device)
    return _LOSS_SCRATCH[key]


class RgbLossFunction(torch.autograd.Function):
    """`sample_grid` + `_rgb_metrics` (pipelines/utils.py:137-158, 272-296) as one launch each way."""

    @staticmethod
    def forward(ctx, pred, image, xy):
        B, n, C = pred.shape
        Hh, Ww = image.shape[1], image.shape[2]
        pred, image, xy = N.f32c(pred), N.f32c(image), N.f32c(xy)
        mse = torch.empty(B, device=pred.device)
        huber = torch.empty(B, device=pred.device)
        _call("yn_rgb_loss_fwd", N.ptr(pred), N.ptr(image), N.ptr(xy), N.ptr(mse), N.ptr(huber),
              N.ptr(_loss_scratch(B, pred.device), torch.uint8), B, n, C, Ww, Hh, STREAM, device=pred.device)
        ctx.save_for_backward(pred, image, xy, mse)
        return mse, huber

    @staticmethod
    def backward(ctx, g_mse, g_huber):
        pred, image, xy, mse = ctx.saved_tensors
        B, n, C = pred.shape
        g_mse = None if g_mse is None else N.f32c(g_mse)
        g_huber = None if g_huber is None else N.f32c(g_huber)
        d_pred = torch.empty_like(pred)
        _call("yn_rgb_loss_bwd", N.ptr(pred), N.ptr(image), N.ptr(xy), N.ptr(mse), N.ptr(g_mse), N.ptr(g_huber), N.ptr(d_pred),
              B, n, C, image.shape[2], image.shape[1], STREAM, device=pred.device)
        return d_pred, None, None


def rgb_loss(pred: torch.Tensor, image: torch.Tensor, xy: torch.Tensor):
    """pred [B,n,C], image [B,H,W,C], xy [B,n,2] float pixel coordinates -> (mse [B], huber [B])."""
    return RgbLossFunction.apply(pred, image, xy)


def scatter_rays(tensors: Sequence[torch.Tensor], xy: torch.Tensor, height: int, width: int):
    """`scatter_rays_to_image` (pipelines/utils.py:299-323) for up to three `[B,n,C_k]` tensors at once: ONE zero fill of
    a shared allocation + ONE launch; returns contiguous canvases `[B,H,W,C_k]`."""
    B, n = xy.shape[0], xy.shape[1]
    dev = xy.device
    srcs = [N.f32c(t.reshape(B, n, t.shape[-1])) for t in tensors]
    chans = [t.shape[-1] for t in srcs]
    buf = torch.zeros(B * height * width * sum(chans), device=dev)
    outs, off = [], 0
    for c in chans:
        k = B * height * width * c
        outs.append(buf[off:off + k].view(B, height, width, c))
        off += k
    xy = N.f32c(xy)
    m = len(srcs)
    src_arr = (ctypes.c_void_p * m)(*[t.data_ptr() for t in srcs])
    dst_arr = (ctypes.c_void_p * m)(*[t.data_ptr() for t in outs])
    ch_arr = (ctypes.c_int * m)(*chans)
    _call("yn_scatter_rays", src_arr, dst_arr, ch_arr, m, N.ptr(xy), B, n, width, height, STREAM, device=dev)
    return outs


# --------------------------------------------------------------------------- #
# optimizer
# --------------------------------------------------------------------------- #
def adam_step(params, grads, exp_avg, exp_avg_sq, lr, step, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    _call("yn_adam_step",
          N.ptr(params), N.ptr(grads), N.ptr(exp_avg), N.ptr(exp_avg_sq), params.numel(), lr, beta1, beta2, eps,
          int(step), grad_scale, STREAM, device=params.device)


def adam_step_dev(params, grads, exp_avg, exp_avg_sq, state, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    """Adam with step / lr read from the device tensor `state` = [step, lr] (CUDA-graph friendly)."""
    _call("yn_adam_step_dev", N.ptr(params), N.ptr(grads), N.ptr(exp_avg), N.ptr(exp_avg_sq), params.numel(),
          N.ptr(state), beta1, beta2, eps, grad_scale, STREAM, device=params.device)
