# flake8: noqa
from .builder import FEATURE_EXTRACTORS
from . import identity_mapper
