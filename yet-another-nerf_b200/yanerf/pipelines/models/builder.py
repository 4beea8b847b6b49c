from yanerf.utils.registry import Registry

MODELS = Registry("models")
