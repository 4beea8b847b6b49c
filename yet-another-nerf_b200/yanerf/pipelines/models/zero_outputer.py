"""`ZeroOutputer`: debug implicit function with zero density and colour
(yanerf/pipelines/models/zero_outputer.py:13-36); used by the known-answer pipeline test."""
import warnings

import torch

from .builder import MODELS


@MODELS.register_module()
class ZeroOutputer(torch.nn.Module):
    def __init__(self) -> None:
        super().__init__()
        warnings.warn("Should not use ZeroOutputer, Debug only.")

    def forward(self, origins, directions, lengths, global_codes=None, **kwargs) -> dict:
        return dict(
            rays_densities=lengths.new_zeros(*lengths.shape, 1),
            rays_features=lengths.new_zeros(*lengths.shape, 3),
            aux={},
        )
