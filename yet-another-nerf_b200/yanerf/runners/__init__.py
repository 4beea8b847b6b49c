from .engine import FusedTrainer, exponential_lr  # noqa: F401
