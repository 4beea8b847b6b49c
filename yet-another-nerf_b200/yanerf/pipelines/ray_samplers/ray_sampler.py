"""RaySampler: pixel pick -> ray bundle (origins, un-normalised directions, depths, xys).

API mirror of `yanerf/pipelines/ray_samplers/ray_sampler.py` (RaySampler 10-115, _RaySampler 118-246,
_safe_multinomial 317-358, get_min_max_depth_bounds 389-401).  The pixel pick stays a torch device op
(`torch.multinomial`); the bundle itself (`_xy_to_ray_bundle` 249-314 and
`_jiggle_within_stratas` 361-386) is one `yn_ray_bundle` launch.
"""
from __future__ import annotations

from typing import List, Optional, Tuple, Union

import torch

from ... import ops

from .builder import RAY_SAMPLERS
from ...pipelines.utils import as_mode
from .utils import EvaluationMode, RayBundle, RenderSamplingMode


def _safe_multinomial(input: torch.Tensor, num_samples: int, all_positive: bool = False) -> torch.Tensor:
    """Without replacement where a row has enough non-zero weights, with replacement otherwise.
    `all_positive`: the caller built `input` as all-ones (no mask), so no device->host check is needed."""
    if all_positive and input.shape[-1] >= num_samples:
        return torch.multinomial(input, num_samples, replacement=False)
    enough = (input > 0.0).sum(dim=-1) >= num_samples
    if bool(enough.all()):
        return torch.multinomial(input, num_samples, replacement=False)
    res = torch.multinomial(input, num_samples, replacement=True)
    if bool(enough.any()):
        res[enough] = torch.multinomial(input[enough], num_samples, replacement=False)
    return res


def get_min_max_depth_bounds(poses, scene_center, scene_extent) -> Tuple[float, float]:
    """near/far = |camera centre - scene centre| -/+ extent, averaged over the batch."""
    cam_center = poses[:, :, -1]
    dist = ((cam_center - poses[:, :3, :-1] @ scene_center) ** 2).sum(dim=-1).clamp(0.001).sqrt()
    dist = dist.clamp(scene_extent + 1e-3)
    return (dist - scene_extent).mean().item(), (dist + scene_extent).mean().item()


_DEPTH_ROWS = {}


def _depth_row(min_depth: float, max_depth: float, n: int, device) -> torch.Tensor:
    """torch.linspace evaluated on the CPU (the parity oracle's device), cached per device."""
    key = (float(min_depth), float(max_depth), n, str(device))
    if key not in _DEPTH_ROWS:
        if len(_DEPTH_ROWS) > 64:
            _DEPTH_ROWS.clear()
        _DEPTH_ROWS[key] = torch.linspace(min_depth, max_depth, n, dtype=torch.float32).to(device)
    return _DEPTH_ROWS[key]


_PIXEL_GRIDS = {}


def _pixel_grid(H: int, W: int, device) -> torch.Tensor:
    """Float pixel coordinates (x, y) of the flattened H x W grid, [H*W, 2], cached per device."""
    key = (H, W, str(device))
    if key not in _PIXEL_GRIDS:
        if len(_PIXEL_GRIDS) > 16:
            _PIXEL_GRIDS.clear()
        idx = torch.arange(H * W, device=device)
        _PIXEL_GRIDS[key] = torch.stack((idx % W, idx // W), dim=-1).float()
    return _PIXEL_GRIDS[key]


class _RaySampler(torch.nn.Module):
    def __init__(self, *, image_width: int, image_height: int, n_pts_per_ray: int, min_depth: float,
                 max_depth: float, n_rays_per_image: Optional[int] = None, unit_directions: bool = False,
                 stratified_sampling: bool = False) -> None:
        super().__init__()
        self._image_width = image_width
        self._image_height = image_height
        self._n_pts_per_ray = n_pts_per_ray
        self._min_depth = min_depth
        self._max_depth = max_depth
        self._n_rays_per_image = n_rays_per_image
        self._unit_directions = unit_directions  # stored, never applied (reference behaviour)
        self._stratified_sampling = stratified_sampling
        # Unmasked training pick through `yn_sample_pixels` instead of torch.multinomial over H*W weights (same
        # distribution: a uniformly random n-subset; O(n) and CUDA-graph friendly).  Off by default: the reference
        # consumes torch's global generator here.  FusedTrainer turns it on.
        self.fused_pixel_sampler = False
        self._pixel_seed: Optional[torch.Tensor] = None
        # ops.DeviceRng: with the fused pixel sampler, pick + rays + stratified jitter become ONE launch (`yn_train_rays`)
        # whose draws are generated in the kernel; None = torch's generators (`torch.rand`), as in the reference
        self.device_rng = None

    def forward(self, poses, focal_lengths, *, image_height=None, image_width=None, mask=None,
                sampling_prob_mask=None, min_depth=None, max_depth=None,
                n_rays_per_image: Union[None, int, List[int]] = None, n_pts_per_ray=None,
                stratified_sampling=None, ray_range: Optional[Tuple[int, int]] = None) -> RayBundle:
        """ray_range=(start, end): full-grid mode only: the rays of flat pixels [start, end) instead of the whole H x W
        grid, spatial shape (end - start, 1) (one rank's slab of a ray-sharded render)."""
        B = poses.shape[0]
        device = poses.device
        poses = poses[:, :3, :4]
        if image_height is None or image_width is None:
            image_height, image_width = self._image_height, self._image_width
        H, W = image_height, image_width
        num_rays = n_rays_per_image or self._n_rays_per_image
        if mask is not None and num_rays is None:
            num_rays = mask.sum(dim=(1, 2)).min().int().item()

        xy = None
        spatial: Tuple[int, ...] = (H, W)
        if num_rays is not None:
            plain = mask is None and sampling_prob_mask is None
            fused = plain and self.fused_pixel_sampler and isinstance(num_rays, int) and poses.is_cuda
            if mask is not None:
                assert tuple(mask.shape) == (B, H, W)
                weights = mask.reshape(B, -1).float()
            elif not fused:
                weights = torch.ones(B, H * W, device=device)
            if sampling_prob_mask is not None:
                if tuple(sampling_prob_mask.shape) == (B, H, W):
                    weights = weights * sampling_prob_mask.reshape(B, -1)
                elif sampling_prob_mask.ndim == 4:
                    if isinstance(num_rays, int):
                        num_rays = [num_rays]
                    if tuple(sampling_prob_mask[:, 0].shape) != (B, H, W):
                        raise ValueError(
                            f"Invalid `sampling_prob_mask`: `sampling_prob_mask.shape` {sampling_prob_mask.shape}, "
                            f"must align with {(B, H, W, 2)}"
                        )
                    if sampling_prob_mask.shape[1] != len(num_rays):
                        raise ValueError(
                            f"Invalid number of sampling layers: sampling_prob_mask.shape[1] "
                            f"{sampling_prob_mask.shape[1]} vs. len(num_rays) {len(num_rays)}"
                        )
                    weights = weights[:, None] * sampling_prob_mask.reshape(B, len(num_rays), -1)
                else:
                    raise ValueError(
                        f"Invalida `sampling_prob_mask`, shape of {sampling_prob_mask.shape}, want (B, H, W) or (B, L, H, W)"
                    )
            xy_fused = None
            if fused and self.device_rng is not None:
                md = self._min_depth if min_depth is None else min_depth
                xd = self._max_depth if max_depth is None else max_depth
                if isinstance(md, torch.Tensor):
                    md = md.mean().item()
                if isinstance(xd, torch.Tensor):
                    xd = xd.mean().item()
                n_pts = self._n_pts_per_ray if n_pts_per_ray is None else n_pts_per_ray
                stratified = self._stratified_sampling if stratified_sampling is None else stratified_sampling
                _, xys, o, d, z = ops.train_rays(self.device_rng, poses, focal_lengths, _depth_row(md, xd, n_pts, device),
                                                 bool(stratified and n_pts > 0), num_rays, self._image_width,
                                                 self._image_height) if (W, H) == (self._image_width, self._image_height) else (None,) * 5
                if xys is not None:
                    sp = (num_rays, 1)
                    return RayBundle(origins=o.reshape(B, *sp, 3), directions=d.reshape(B, *sp, 3),
                                     lengths=z.reshape(B, *sp, n_pts), xys=xys.reshape(B, *sp, 2))
            if fused:
                if self._pixel_seed is None or self._pixel_seed.device != device:
                    # seeded from torch's CPU generator: reproducible under torch.manual_seed, no device sync
                    self._pixel_seed = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).to(device)
                rays_idx, xy_fused = ops.sample_pixels(self._pixel_seed, B, num_rays, W, H)
                self._pixel_seed += 1
            elif weights.ndim == 2:
                rays_idx = _safe_multinomial(weights, num_rays, all_positive=plain)
            else:
                rays_idx = torch.cat([_safe_multinomial(weights[:, i], num_rays[i]) for i in range(len(num_rays))], dim=-1)
            xy = xy_fused if xy_fused is not None else torch.stack((rays_idx % W, rays_idx // W), dim=-1).float()
            spatial = (rays_idx.shape[1], 1)

        min_depth = self._min_depth if min_depth is None else min_depth
        max_depth = self._max_depth if max_depth is None else max_depth
        if isinstance(min_depth, torch.Tensor):
            min_depth = min_depth.mean().item()
        if isinstance(max_depth, torch.Tensor):
            max_depth = max_depth.mean().item()
        n_pts = self._n_pts_per_ray if n_pts_per_ray is None else n_pts_per_ray
        stratified = self._stratified_sampling if stratified_sampling is None else stratified_sampling

        if ray_range is not None and xy is None:
            start, end = ray_range
            xy = _pixel_grid(H, W, device)[start:end][None].expand(B, -1, -1).contiguous()
            spatial = (end - start, 1)
        n = spatial[0] * spatial[1]
        depths = _depth_row(min_depth, max_depth, n_pts, device)
        u = torch.rand(B, n, n_pts, device=device) if (stratified and n_pts > 0) else None
        # NB: the pinhole uses the sampler's CONFIGURED width/height even for custom image sizes
        # (ray_sampler.py:238-239, 302-303); the pixel grid uses the custom ones.
        o, d, z, xys = ops.ray_bundle(
            poses, focal_lengths, xy, depths, u, n, self._image_width, self._image_height
        ) if xy is not None else _full_grid_bundle(poses, focal_lengths, depths, u, H, W, self._image_width, self._image_height)
        return RayBundle(
            origins=o.reshape(B, *spatial, 3), directions=d.reshape(B, *spatial, 3),
            lengths=z.reshape(B, *spatial, n_pts), xys=xys.reshape(B, *spatial, 2),
        )


def _full_grid_bundle(poses, focal, depths, u, H, W, cfg_w, cfg_h):
    if (W, H) == (cfg_w, cfg_h):
        return ops.ray_bundle(poses, focal, None, depths, u, H * W, cfg_w, cfg_h)
    # custom image size: explicit pixel list of the H x W grid, pinhole centred on the configured size
    B = poses.shape[0]
    idx = torch.arange(H * W, device=poses.device)
    xy = torch.stack((idx % W, idx // W), dim=-1).float()[None].expand(B, -1, -1).contiguous()
    return ops.ray_bundle(poses, focal, xy, depths, u, H * W, cfg_w, cfg_h)


@RAY_SAMPLERS.register_module()
class RaySampler(torch.nn.Module):
    def __init__(self, image_width: int = 400, image_height: int = 400,
                 scene_center: Tuple[float, float, float] = (0.0, 0.0, 0.0), scene_extent: float = 0.0,
                 sampling_mode_training: str = "mask_sample", sampling_mode_evaluation: str = "full_grid",
                 n_pts_per_ray_training: int = 64, n_pts_per_ray_evaluation: int = 64,
                 n_rays_per_image_sampled_from_mask: int = 1024, min_depth: float = 0.1, max_depth: float = 8.0,
                 stratified_point_sampling_training: bool = True,
                 stratified_point_sampling_evaluation: bool = False) -> None:
        super().__init__()
        self.image_width, self.image_height = image_width, image_height
        self._sampling_mode = {
            EvaluationMode.TRAINING: RenderSamplingMode(sampling_mode_training),
            EvaluationMode.EVALUATION: RenderSamplingMode(sampling_mode_evaluation),
        }

        def make(mode, n_pts, stratified):
            masked = self._sampling_mode[mode] == RenderSamplingMode.MASK_SAMPLE
            return _RaySampler(
                image_width=image_width, image_height=image_height, n_pts_per_ray=n_pts, min_depth=min_depth,
                max_depth=max_depth, n_rays_per_image=n_rays_per_image_sampled_from_mask if masked else None,
                unit_directions=True, stratified_sampling=stratified,
            )

        self._raysamplers = {
            EvaluationMode.TRAINING: make(EvaluationMode.TRAINING, n_pts_per_ray_training, stratified_point_sampling_training),
            EvaluationMode.EVALUATION: make(EvaluationMode.EVALUATION, n_pts_per_ray_evaluation, stratified_point_sampling_evaluation),
        }
        self.register_buffer("scene_center", torch.tensor(scene_center, dtype=torch.float32), persistent=False)
        self.scene_extent = scene_extent

    @property
    def fused_pixel_sampler(self) -> bool:
        """Unmasked pixel picks through `yn_sample_pixels` instead of torch.multinomial (see _RaySampler)."""
        return all(s.fused_pixel_sampler for s in self._raysamplers.values())

    @fused_pixel_sampler.setter
    def fused_pixel_sampler(self, value: bool) -> None:
        for s in self._raysamplers.values():  # plain dict, not sub-modules: `.modules()` does not reach them
            s.fused_pixel_sampler = bool(value)

    def set_device_rng(self, rng) -> None:
        """In-kernel draws for the training sampler (ops.DeviceRng; None switches back to torch's generators)."""
        for s in self._raysamplers.values():
            s.device_rng = rng

    def forward(self, poses, focal_lengths, evaluation_mode: EvaluationMode, *, mask=None, sampling_prob_mask=None,
                image_height=None, image_width=None, min_depth=None, max_depth=None,
                n_rays_per_image: Union[None, int, List[int]] = None, ray_range: Optional[Tuple[int, int]] = None) -> RayBundle:
        evaluation_mode = as_mode(evaluation_mode)
        sample_mask = None
        if self._sampling_mode[evaluation_mode] == RenderSamplingMode.MASK_SAMPLE and mask is not None:
            h = self.image_height if image_height is None or image_width is None else image_height
            w = self.image_width if image_height is None or image_width is None else image_width
            sample_mask = torch.nn.functional.interpolate(mask, size=[h, w], mode="nearest")[:, 0]
        if min_depth is None and max_depth is None and self.scene_extent > 0.0:
            min_depth, max_depth = get_min_max_depth_bounds(poses, self.scene_center, self.scene_extent)
        return self._raysamplers[evaluation_mode](
            poses, focal_lengths, mask=sample_mask, sampling_prob_mask=sampling_prob_mask, min_depth=min_depth,
            max_depth=max_depth, n_rays_per_image=n_rays_per_image, image_height=image_height, image_width=image_width,
            ray_range=ray_range,
        )
