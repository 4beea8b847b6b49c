"""Import shims that let the UNMODIFIED reference (/root/reference, torch 1.10 era)
import under this image's Python 3.12 / torch 2.11.  Used only by
make_golden.py in the build container; never on the GPU box, never by the product.
"""
import sys
import types


def install() -> None:
    # addict.Dict: attribute dict with recursive wrapping (utils/config.py:18)
    class Dict(dict):
        def __init__(self, *args, **kwargs):
            super().__init__()
            for k, v in dict(*args, **kwargs).items():
                self[k] = self._hook(v)

        @classmethod
        def _hook(cls, v):
            if isinstance(v, dict) and not isinstance(v, cls):
                return cls(v)
            if isinstance(v, (list, tuple)):
                return type(v)(cls._hook(i) for i in v)
            return v

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                return self.__missing__(k)

        def __missing__(self, k):
            raise KeyError(k)

        def __setattr__(self, k, v):
            self[k] = self._hook(v)

        def __setitem__(self, k, v):
            super().__setitem__(k, self._hook(v))

        def to_dict(self):
            out = {}
            for k, v in self.items():
                if isinstance(v, Dict):
                    out[k] = v.to_dict()
                elif isinstance(v, (list, tuple)):
                    out[k] = type(v)(i.to_dict() if isinstance(i, Dict) else i for i in v)
                else:
                    out[k] = v
            return out

    addict = types.ModuleType("addict")
    addict.Dict = Dict
    sys.modules["addict"] = addict

    yapf = types.ModuleType("yapf")
    yapflib = types.ModuleType("yapf.yapflib")
    yapf_api = types.ModuleType("yapf.yapflib.yapf_api")
    yapf_api.FormatCode = lambda text, **kw: (text, False)
    yapf.yapflib = yapflib
    yapflib.yapf_api = yapf_api
    sys.modules.update({"yapf": yapf, "yapf.yapflib": yapflib, "yapf.yapflib.yapf_api": yapf_api})

    imageio = types.ModuleType("imageio")

    def imread(path):
        import numpy as np
        from PIL import Image

        im = Image.open(path)
        if im.mode == "P":
            im = im.convert("RGBA")
        return np.asarray(im)

    def imwrite(path, arr):
        from PIL import Image

        Image.fromarray(arr).save(path)

    imageio.imread, imageio.imwrite = imread, imwrite
    sys.modules["imageio"] = imageio

    omegaconf = types.ModuleType("omegaconf")
    omegaconf.DictConfig = type("DictConfig", (dict,), {})
    sys.modules["omegaconf"] = omegaconf

    six = types.ModuleType("torch._six")
    six.string_classes = (str, bytes)
    sys.modules["torch._six"] = six
