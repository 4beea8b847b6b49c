# flake8: noqa
from .builder import PIPELINES
from . import nerf_pipeline
