from .engine import FusedTrainer, reference_lr, scaled_runner_config  # noqa: F401
from .device_feed import DeviceSceneFeed  # noqa: F401
