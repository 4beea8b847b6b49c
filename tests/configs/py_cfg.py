from yanerf.utils.config import Config

_base = Config.fromfile("{{ fileDirname }}/base.yml")
pipeline = dict(**_base.pipeline)
pipeline["num_passes"] = 3
