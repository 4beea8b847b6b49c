from yanerf.utils.registry import Registry

PIPELINES = Registry("pipelines")
