# flake8: noqa
from .builder import RENDERERS
from . import multipass_emission_absorpsion_renderer
