"""`IdentityMapper` (yanerf/pipelines/feature_extractors/identity_mapper.py:5-11): returns its kwargs."""
import torch

from .builder import FEATURE_EXTRACTORS


@FEATURE_EXTRACTORS.register_module()
class IdentityMapper(torch.nn.Module):
    def forward(self, **kwargs):
        return kwargs
