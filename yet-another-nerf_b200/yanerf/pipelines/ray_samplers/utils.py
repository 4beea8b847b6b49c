"""Ray sampler helpers (yanerf/pipelines/ray_samplers/utils.py:7-24)."""
from enum import Enum

import torch

from ...pipelines.utils import EvaluationMode, RayBundle  # noqa: F401  (re-exported like the reference)


class RenderSamplingMode(Enum):
    MASK_SAMPLE = "mask_sample"
    FULL_GRID = "full_grid"


def get_xy_grid(image_height: int, image_width: int) -> torch.Tensor:
    """[H, W, 2] float pixel coordinates stacked (x, y)."""
    ys = torch.arange(image_height, dtype=torch.float32)
    xs = torch.arange(image_width, dtype=torch.float32)
    return torch.stack((xs[None, :].expand(image_height, -1), ys[:, None].expand(-1, image_width)), dim=-1)
