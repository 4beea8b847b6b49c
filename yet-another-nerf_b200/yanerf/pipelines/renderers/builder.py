from yanerf.utils.registry import Registry

RENDERERS = Registry("renderers")
