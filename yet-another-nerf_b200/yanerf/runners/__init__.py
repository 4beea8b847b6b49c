from .engine import FusedTrainer, exponential_lr  # noqa: F401
from .device_feed import DeviceSceneFeed  # noqa: F401
