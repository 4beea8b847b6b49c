"""Shared pipeline types, per-sample losses and pixel gather/scatter.

API mirror of the reference's `yanerf/pipelines/utils.py` (EvaluationMode 8-10, RayBundle 13-17,
PartialFunctionWrapper 20-33, ViewMetrics 36-134, _rgb_metrics 137-158, huber 189-203,
sample_grid 272-296, scatter_rays_to_image 299-323).  These are a handful of tiny
reductions/gathers per step; they run as torch device ops.
"""
from __future__ import annotations

from enum import Enum
from typing import Any, Dict, NamedTuple, Optional

import torch


class EvaluationMode(Enum):
    TRAINING = "training"
    EVALUATION = "evaluation"


def as_mode(mode) -> "EvaluationMode":
    """Accepts this package's enum, its value string, or the enum of ANOTHER copy of the package (the
    reference's `yanerf.pipelines.utils.EvaluationMode` when our classes are registered into the reference's registries)."""
    return mode if isinstance(mode, EvaluationMode) else EvaluationMode(getattr(mode, "value", mode))


class RayBundle(NamedTuple):
    origins: torch.Tensor
    directions: torch.Tensor
    lengths: torch.Tensor
    xys: torch.Tensor


class PartialFunctionWrapper(torch.nn.Module):
    """Holds an implicit function plus per-call bound keyword arguments (not re-entrant)."""

    def __init__(self, fn: torch.nn.Module):
        super().__init__()
        self._fn = fn
        self.bound_args: Dict[str, Any] = {}

    def bind_args(self, **bound_args):
        self.bound_args = bound_args

    def unbind_args(self):
        self.bound_args = {}

    def forward(self, *args, **kwargs):
        return self._fn(*args, **{**kwargs, **self.bound_args})


def _flat_pixel_index(grid: torch.Tensor, width: int) -> torch.Tensor:
    flat = grid.reshape(grid.shape[0], -1, 2)
    return (flat[:, :, 0] + width * flat[:, :, 1]).long()


def sample_grid(tensor: torch.Tensor, image_sampling_grid: torch.Tensor, validate: bool = True) -> torch.Tensor:
    """Gather `tensor` [B,H,W,C] at float pixel coordinates `image_sampling_grid` [B,*sp,2] -> [B,*sp,C].
    `validate` keeps the reference's range assertions (two device->host syncs); callers that produced the grid
    themselves may skip them."""
    B, H, W, C = tensor.shape
    if validate:
        assert image_sampling_grid[..., 0].max() < W, "Invalid ray_sampler.image_width"
        assert image_sampling_grid[..., 1].max() < H, "Invalid ray_sampler.image_height"
    idx = _flat_pixel_index(image_sampling_grid, W)[:, :, None].expand(-1, -1, C)
    out = torch.gather(tensor.reshape(B, H * W, C), 1, idx)
    return out.reshape(B, *image_sampling_grid.shape[1:-1], C)


@torch.no_grad()
def scatter_rays_to_image(tensor, image_sampling_grid, image_height: int, image_width: int, bg_color=None):
    """Splat per-ray values [B,*sp,C] onto a zero (or bg-coloured) [B,H,W,C] canvas."""
    B, C = tensor.shape[0], tensor.shape[-1]
    assert list(tensor.shape[1:-1]) == list(image_sampling_grid.shape[1:-1]), (
        f"{list(tensor.shape[1:-1])} vs. {list(image_sampling_grid.shape[1:-1])}"
    )
    canvas = tensor.new_zeros(B, image_height, image_width, C)
    if bg_color is not None and bg_color.shape[-1] == C:
        canvas = canvas + bg_color
    canvas = canvas.reshape(B, -1, C)
    idx = _flat_pixel_index(image_sampling_grid, image_width)[:, :, None].expand(-1, -1, C)
    canvas.scatter_(1, idx, tensor.reshape(B, -1, C))
    return canvas.reshape(B, image_height, image_width, C)


def safe_sqrt(A: torch.Tensor, eps: float = 1e-4) -> torch.Tensor:
    return (torch.clamp(A, 0.0) + eps).sqrt()


def huber(dfsq: torch.Tensor, scaling: float = 0.03) -> torch.Tensor:
    return (safe_sqrt(1 + dfsq / (scaling * scaling), eps=1e-4) - 1) * scaling


def _rgb_metrics(images, images_pred, loss_reweight_masks=None):
    B = images.shape[0]
    diff = (images_pred.reshape(B, -1) - images.reshape(B, -1)) ** 2
    if loss_reweight_masks is not None:
        diff = diff * loss_reweight_masks.reshape(images.shape).reshape(B, -1)
    mse = diff.mean(dim=-1)
    return {"rgb_huber": huber(mse, scaling=0.03), "rgb_mse": mse}


def eval_depth(pred, gt, crop: int = 0, mask=None, get_best_scale: bool = True, mask_thr: float = 0.5,
               best_scale_clamp_thr: float = 1e-4):
    """Depth MSE / abs error with optional best-scale alignment (pipelines/utils.py:206-262)."""
    if mask is None:
        mask = torch.ones_like(gt)
    dmask = (gt > 0.0).float() * (mask > mask_thr).float()
    if crop > 0:
        dmask[..., :crop, :, :] = 0
        dmask[..., -crop:, :, :] = 0
        dmask[..., :, :crop, :] = 0
        dmask[..., :, -crop:, :] = 0
    dims = tuple(range(1, pred.ndim))
    if get_best_scale:
        xy = (pred * gt * dmask).mean(dims)
        xx = (pred * pred * dmask).mean(dims)
        pred = pred * (xy / torch.clamp(xx, best_scale_clamp_thr)).reshape(-1, *([1] * (pred.ndim - 1)))
    df = gt - pred
    mass = torch.clamp(dmask.sum(dims), 1e-4)
    mse_depth = (dmask * (df ** 2)).sum(dims) / mass
    abs_depth = (dmask * df.abs()).sum(dims) / mass
    return mse_depth, abs_depth


class ViewMetrics(torch.nn.Module):
    """Per-image (shape `(B,)`) rgb mse / huber (+ optional depth abs error), keys prefixed."""

    # ray counts per image up to this bound take the fused gather + loss kernel (`yn_rgb_loss_fwd`); None = no bound
    fused_max_rays: Optional[int] = None

    def forward(self, image_sampling_grid, images=None, images_pred=None, depths=None, depths_pred=None,
                loss_reweight_masks=None, keys_prefix: Optional[str] = "loss_", validate_grid: bool = True):
        pick = lambda t: None if t is None else sample_grid(t, image_sampling_grid, validate_grid)
        preds = {}
        fused = (images is not None and images_pred is not None and loss_reweight_masks is None and images.is_cuda
                 and images.ndim == 4 and images.shape[-1] == images_pred.shape[-1]
                 and image_sampling_grid.numel() > 0
                 and (self.fused_max_rays is None
                      or image_sampling_grid.numel() // (2 * images.shape[0]) <= self.fused_max_rays))
        if fused:
            # ground-truth gather + squared-error mean in one launch (and one for the backward): `yn_rgb_loss_fwd`
            from .. import ops

            if validate_grid:  # the reference's range assertions (two device->host syncs)
                assert image_sampling_grid[..., 0].max() < images.shape[2], "Invalid ray_sampler.image_width"
                assert image_sampling_grid[..., 1].max() < images.shape[1], "Invalid ray_sampler.image_height"
            B = images.shape[0]
            mse, hub = ops.rgb_loss(images_pred.reshape(B, -1, images_pred.shape[-1]), images,
                                    image_sampling_grid.reshape(B, -1, 2))
            preds.update({"rgb_huber": hub, "rgb_mse": mse})
            images = None
        images, depths, loss_reweight_masks = pick(images), pick(depths), pick(loss_reweight_masks)
        if images is not None and images_pred is not None:
            preds.update(_rgb_metrics(images, images_pred, loss_reweight_masks))
        if depths is not None and depths_pred is not None:
            _, abs_ = eval_depth(depths_pred, depths, get_best_scale=True, mask=None, crop=0)
            preds["depth_abs"] = abs_.mean(dim=-1)
        if keys_prefix is not None:
            preds = {keys_prefix + k: v for k, v in preds.items()}
        return preds
