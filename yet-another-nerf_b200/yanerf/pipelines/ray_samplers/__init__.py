# flake8: noqa
from .builder import RAY_SAMPLERS
from . import ray_sampler
