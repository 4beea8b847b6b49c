"""INTEGRATION.md option B: keep the reference tree, override its hot-path classes through its own registries.

Put this file (or the repository root) on PYTHONPATH next to the reference and add to a reference config

    custom_imports:
      imports: [yanerf_b200_plugin]

`Config.fromfile` imports the module when the config is loaded (`yanerf/utils/config.py:320-324`); the module loads this
repository's package under the private name `yanerf_b200` (its modules import each other relatively, so they never touch
the reference's `yanerf`), and registers its classes behind the reference's `type:` names with
`register_module(force=True)` (`yanerf/utils/registry.py:234-238, 252-305`).  `PIPELINES.build(cfg.pipeline)` of the
reference then returns this repository's `NeRFPipeline`, whose ray sampler, NeRF MLPs and renderer are the sm_100a kernels;
`scripts/run.py` is unchanged.  The reference's `EvaluationMode` enum is accepted at every forward (`as_mode`).
"""
import importlib.util
import os
import sys

_ROOT = os.environ.get("YANERF_B200_ROOT", os.path.dirname(os.path.abspath(__file__)))
_PKG = os.path.join(_ROOT, "yet-another-nerf_b200", "yanerf")
ALIAS = "yanerf_b200"


def _load_alias():
    if ALIAS in sys.modules:
        return sys.modules[ALIAS]
    spec = importlib.util.spec_from_file_location(ALIAS, os.path.join(_PKG, "__init__.py"), submodule_search_locations=[_PKG])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[ALIAS] = mod
    spec.loader.exec_module(mod)
    return mod


_load_alias()
import yanerf_b200.pipelines as _b200_pipelines  # noqa: E402  (registers the classes in the alias package's own registries)
from yanerf_b200.pipelines.feature_extractors.identity_mapper import IdentityMapper  # noqa: E402
from yanerf_b200.pipelines.models.nerf_mlp import NeRFMLP  # noqa: E402
from yanerf_b200.pipelines.models.zero_outputer import ZeroOutputer  # noqa: E402
from yanerf_b200.pipelines.nerf_pipeline import NeRFPipeline  # noqa: E402
from yanerf_b200.pipelines.ray_samplers.ray_sampler import RaySampler  # noqa: E402
from yanerf_b200.pipelines.renderers.multipass_emission_absorpsion_renderer import MultipassEmissionAbsorpsionRenderer  # noqa: E402

OVERRIDES = {
    "PIPELINES": {"NeRFPipeline": NeRFPipeline},
    "MODELS": {"NeRFMLP": NeRFMLP, "ZeroOutputer": ZeroOutputer},
    "RENDERERS": {"MultipassEmissionAbsorpsionRenderer": MultipassEmissionAbsorpsionRenderer},
    "RAY_SAMPLERS": {"RaySampler": RaySampler},
    "FEATURE_EXTRACTORS": {"IdentityMapper": IdentityMapper},
}


def install() -> None:
    """Register this repository's classes into the registries of whatever `yanerf` is importable (the reference's)."""
    import yanerf.pipelines as ref_pipelines  # noqa: F401  (the reference registers its own classes first)
    from yanerf.pipelines.builder import PIPELINES
    from yanerf.pipelines.feature_extractors.builder import FEATURE_EXTRACTORS
    from yanerf.pipelines.models.builder import MODELS
    from yanerf.pipelines.ray_samplers.builder import RAY_SAMPLERS
    from yanerf.pipelines.renderers.builder import RENDERERS

    regs = dict(PIPELINES=PIPELINES, MODELS=MODELS, RENDERERS=RENDERERS, RAY_SAMPLERS=RAY_SAMPLERS,
                FEATURE_EXTRACTORS=FEATURE_EXTRACTORS)
    for reg_name, classes in OVERRIDES.items():
        for type_name, cls in classes.items():
            regs[reg_name].register_module(name=type_name, force=True, module=cls)


install()
