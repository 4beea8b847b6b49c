"""Multi-pass emission-absorption renderer: raymarch -> refine -> raymarch ...

API mirror of `yanerf/pipelines/renderers/multipass_emission_absorpsion_renderer.py`
(MultipassEmissionAbsorpsionRenderer 11-117, EmissionAbsorptionRaymarcher 120-239, input checks 242-278);
compositing is `yn_composite_fwd` / `yn_composite_bwd`.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Tuple, Union

import torch

from ... import ops
from ...pipelines.utils import EvaluationMode, RayBundle, as_mode

from .builder import RENDERERS
from .utils import RayPointRefiner, RendererOutput


def _check_raymarcher_inputs(rays_densities, rays_features, rays_z, features_can_be_none=False,
                             z_can_be_none=False, density_1d=True) -> None:
    if not torch.is_tensor(rays_densities):
        raise ValueError("rays_densities has to be an instance of torch.Tensor.")
    if not z_can_be_none and not torch.is_tensor(rays_z):
        raise ValueError("rays_z has to be an instance of torch.Tensor.")
    if not features_can_be_none and not torch.is_tensor(rays_features):
        raise ValueError("rays_features has to be an instance of torch.Tensor.")
    if rays_densities.ndim < 1:
        raise ValueError("rays_densities have to have at least one dimension.")
    if density_1d and rays_densities.shape[-1] != 1:
        raise ValueError("The size of the last dimension of rays_densities has to be one.")
    rays_shape = rays_densities.shape[:-1]
    if not z_can_be_none and rays_z.shape != rays_shape:
        raise ValueError("rays_z have to be of the same shape as rays_densities.")
    if not features_can_be_none and rays_features.shape[:-1] != rays_shape:
        raise ValueError(
            "The first to previous to last dimensions of rays_features"
            " have to be the same as all dimensions of rays_densities."
        )


class EmissionAbsorptionRaymarcher(torch.nn.Module):
    def __init__(self, surface_thickness: int = 1, bg_color: Union[Tuple[float, ...], torch.Tensor] = (0.0,),
                 capping_function: str = "exponential", weight_function: str = "product",
                 background_opacity: float = 1e10, density_relu: bool = True, blend_output: bool = True,
                 background_density_bias: float = 0.0, hard_background: bool = False) -> None:
        super().__init__()
        if capping_function not in ("exponential", "cap1"):
            raise KeyError(capping_function)
        if weight_function not in ("product", "minimum"):
            raise KeyError(weight_function)
        if capping_function != "exponential" or weight_function != "product" or surface_thickness != 1 or not density_relu:
            raise NotImplementedError(
                "the compositing kernel implements capping_function='exponential', weight_function='product', "
                "surface_thickness=1, density_relu=True (what MultipassEmissionAbsorpsionRenderer builds)"
            )
        self.surface_thickness = surface_thickness
        self.density_relu = density_relu
        self.background_opacity = background_opacity
        self.blend_output = blend_output
        self.background_density_bias = background_density_bias
        self.hard_background = hard_background
        if not isinstance(bg_color, torch.Tensor):
            bg_color = torch.tensor(bg_color, dtype=torch.float32)
        self.register_buffer("_bg_color", bg_color, persistent=False)
        self._bg_host = [float(v) for v in bg_color.reshape(-1).tolist()]
        self.device_rng = None  # ops.DeviceRng: density noise drawn in the kernel instead of torch.randn_like

    def forward(self, rays_densities, rays_features, aux: Dict[str, Any], ray_lengths, ray_directions,
                density_noise_std: float = 0.0, bg_color: Optional[torch.Tensor] = None, rng_pass: int = 0):
        """rays_densities `[...,P,1]`, rays_features `[...,P,C]`, ray_lengths `[...,P]`, ray_directions `[...,3]`
        -> features `[...,C]`, depths `[...,1]`, opacities `[...,1]`, weights `[...,P]`, aux."""
        _check_raymarcher_inputs(rays_densities, rays_features, ray_lengths, z_can_be_none=True,
                                 features_can_be_none=False, density_1d=True)
        lead, P, C = ray_lengths.shape[:-1], ray_lengths.shape[-1], rays_features.shape[-1]
        n_bg = bg_color.shape[-1] if bg_color is not None else len(self._bg_host)
        if not self.hard_background and n_bg not in (1, C):
            shape = tuple(bg_color.shape) if bg_color is not None else (n_bg,)
            raise ValueError(f"Wrong number of background color channels: _bg_color {shape} vs. features {(*lead, C)}.")
        if C > 4:
            raise NotImplementedError("the compositing kernel supports up to 4 feature channels")
        cfg = ops.march_cfg(self.background_opacity, self.background_density_bias, float(density_noise_std),
                            self.blend_output, self.hard_background, self._bg_host)
        sigma = rays_densities.reshape(-1, P)
        rng = self.device_rng if density_noise_std > 0.0 else None
        noise = torch.randn_like(sigma) if (density_noise_std > 0.0 and rng is None) else None
        bg = None if bg_color is None else bg_color.expand(*lead, n_bg).reshape(-1, n_bg)
        feats, depths, opac, weights = ops.composite(
            sigma, rays_features.reshape(-1, P, C), ray_lengths.reshape(-1, P),
            ray_directions.expand(*lead, 3).reshape(-1, 3), cfg, noise, bg, rng, ops.DeviceRng.SITE_NOISE + rng_pass,
        )
        return (feats.reshape(*lead, C), depths.reshape(*lead, 1), opac.reshape(*lead, 1),
                weights.reshape(*lead, P), aux)


@RENDERERS.register_module()
class MultipassEmissionAbsorpsionRenderer(torch.nn.Module):
    def __init__(self, n_pts_per_ray_fine_training: int = 64, n_pts_per_ray_fine_evaluation: int = 64,
                 stratified_sampling_coarse_training: bool = True, stratified_sampling_coarse_evaluation: bool = False,
                 append_coarse_samples_to_fine: bool = True, bg_color: Tuple[float, ...] = (0.0,),
                 density_noise_std_train: float = 0.0, capping_function: str = "exponential",
                 weight_function: str = "product", background_opacity: float = 1e10, blend_output: bool = False,
                 background_density_bias: float = 0.0, hard_background: bool = False) -> None:
        super().__init__()
        self.density_noise_std_train = density_noise_std_train
        self._refiners = {
            EvaluationMode.TRAINING: RayPointRefiner(
                n_pts_per_ray=n_pts_per_ray_fine_training, random_sampling=stratified_sampling_coarse_training,
                add_input_samples=append_coarse_samples_to_fine),
            EvaluationMode.EVALUATION: RayPointRefiner(
                n_pts_per_ray=n_pts_per_ray_fine_evaluation, random_sampling=stratified_sampling_coarse_evaluation,
                add_input_samples=append_coarse_samples_to_fine),
        }
        self._raymarcher: Callable = EmissionAbsorptionRaymarcher(
            surface_thickness=1, bg_color=bg_color, capping_function=capping_function,
            weight_function=weight_function, background_opacity=background_opacity, blend_output=blend_output,
            hard_background=hard_background, background_density_bias=background_density_bias)

    def forward(self, origins, directions, lengths, xys, bg_color: Optional[torch.Tensor], *,
                implicit_functions: List[Callable], evaluation_mode: EvaluationMode = EvaluationMode.EVALUATION,
                **kwargs) -> RendererOutput:
        if not implicit_functions:
            raise ValueError("EA renderer expects implicit functions")
        return self._run_raymarcher(origins, directions, lengths, xys, bg_color, implicit_functions, None,
                                    as_mode(evaluation_mode), **kwargs)

    def set_device_rng(self, rng) -> None:
        """In-kernel draws (ops.DeviceRng) for the density noise and the refiners' inverse-CDF uniforms; None switches
        back to torch's generators."""
        self._raymarcher.device_rng = rng
        for r in self._refiners.values():
            r.device_rng = rng

    def _run_raymarcher(self, origins, directions, lengths, xys, bg_color, implicit_functions, prev_stage,
                        evaluation_mode, **kwargs) -> RendererOutput:
        noise_std = self.density_noise_std_train if evaluation_mode == EvaluationMode.TRAINING else 0.0
        pass_idx = 0
        stage = prev_stage
        while stage is not None:
            pass_idx, stage = pass_idx + 1, stage.prev_stage
        features, depths, alpha_masks, weights, aux = self._raymarcher(
            **implicit_functions[0](origins, directions, lengths, **kwargs),
            ray_lengths=lengths, ray_directions=directions, density_noise_std=noise_std, bg_color=bg_color,
            rng_pass=pass_idx,
        )
        aux["weights"] = weights
        output = RendererOutput(features=features, depths=depths, alpha_masks=alpha_masks, aux=aux, prev_stage=prev_stage)
        if len(implicit_functions) > 1:
            bundle: RayBundle = self._refiners[evaluation_mode](origins, directions, lengths, xys, weights, rng_pass=pass_idx)
            output = self._run_raymarcher(*bundle, bg_color, implicit_functions[1:], output, evaluation_mode, **kwargs)
        return output
