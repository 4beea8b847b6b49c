# flake8: noqa
from .builder import MODELS
from . import nerf_mlp, zero_outputer
