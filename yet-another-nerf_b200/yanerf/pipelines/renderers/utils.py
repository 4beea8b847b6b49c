"""Renderer data types, the fine-sample refiner and `sample_pdf`.

API mirror of `yanerf/pipelines/renderers/utils.py` (RendererOutput 11-33, RayPointRefiner 36-69,
sample_pdf / sample_pdf_python 72-158); the arithmetic is `yn_sample_pdf_merge`.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, Optional

import torch

from ... import ops
from ...pipelines.utils import RayBundle


@dataclass
class RendererOutput:
    """features `[B,*sp,C]`, depths `[B,*sp,1]`, alpha_masks `[B,*sp,1]`; `prev_stage` chains coarse passes."""

    features: torch.Tensor
    depths: torch.Tensor
    alpha_masks: torch.Tensor
    prev_stage: Optional["RendererOutput"] = None
    normals: Optional[torch.Tensor] = None
    points: Optional[torch.Tensor] = None
    aux: Dict[str, Any] = field(default_factory=lambda: {})


def _raise_if_flagged(flag: torch.Tensor) -> None:
    # the reference syncs here as well (`if weights.min() <= 0`, renderers/utils.py:123-124)
    if int(flag.item()) != 0:
        raise ValueError("Negative weights provided.")


class RayPointRefiner(torch.nn.Module):
    """Importance-resamples `n_pts_per_ray` depths from the coarse weights, merges and sorts."""

    # set False to skip the device->host read of the "negative weights" flag (fused runners check it later)
    check_weights: bool = True

    def __init__(self, n_pts_per_ray: int, random_sampling: bool, add_input_samples: bool = True) -> None:
        super().__init__()
        self.n_pts_per_ray = n_pts_per_ray
        self.random_sampling = random_sampling
        self.add_input_samples = add_input_samples
        self.last_flag: Optional[torch.Tensor] = None
        self.device_rng = None  # ops.DeviceRng: uniforms drawn in the kernel instead of torch.rand

    def forward(self, origins, directions, lengths, xys, ray_weights, rng_pass: int = 0) -> RayBundle:
        with torch.no_grad():
            lead, P = lengths.shape[:-1], lengths.shape[-1]
            z = lengths.reshape(-1, P)
            w = ray_weights.reshape(-1, P)
            rng = self.device_rng if self.random_sampling else None
            u = torch.rand(z.shape[0], self.n_pts_per_ray, device=z.device) if (self.random_sampling and rng is None) else None
            # the kernel takes the full weight row and uses weights[..., 1:-1] like the reference
            z_new, _, flag = ops.sample_pdf_merge(z, w, self.n_pts_per_ray, u, self.add_input_samples, rng=rng,
                                                  site=ops.DeviceRng.SITE_PDF + rng_pass)
            self.last_flag = flag
            if self.check_weights:
                _raise_if_flagged(flag)
        return RayBundle(origins=origins, directions=directions, lengths=z_new.reshape(*lead, -1), xys=xys)


def sample_pdf(bins: torch.Tensor, weights: torch.Tensor, N_samples: int, det: bool = False, eps: float = 1e-5):
    """`sample_pdf(bins [R,nb], weights [R,nb-1], N_samples, det)` -> `[R,N_samples]` in draw order
    (renderers/utils.py:72-158); raises ValueError on non-positive `weights + eps` like the reference."""
    if eps != 1e-5:
        raise NotImplementedError("the kernel fixes eps = 1e-5 (the only value the reference uses)")
    u = None if det else torch.rand(bins.shape[0], N_samples, device=bins.device)
    samples, _, flag = ops.sample_pdf(bins, weights, N_samples, u)
    _raise_if_flagged(flag)
    return samples


sample_pdf_python = sample_pdf
