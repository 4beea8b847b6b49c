"""Rank-aware logger (file handler on rank 0 only), after `yanerf/utils/logging.py:9-81`."""
import logging

_initialized = {}


def get_logger(name="yanerf", log_file=None, log_level=logging.INFO, file_mode="w"):
    logger = logging.getLogger(name)
    if name in _initialized:
        return logger
    for known in _initialized:
        if name.startswith(known):
            return logger
    try:
        import torch.distributed as dist

        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    except Exception:
        rank = 0
    handlers = [logging.StreamHandler()]
    if rank == 0 and log_file is not None:
        handlers.append(logging.FileHandler(log_file, file_mode))
    fmt = logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s")
    for h in handlers:
        h.setFormatter(fmt)
        h.setLevel(log_level)
        logger.addHandler(h)
    logger.setLevel(log_level if rank == 0 else logging.ERROR)
    logger.propagate = False
    _initialized[name] = True
    return logger
