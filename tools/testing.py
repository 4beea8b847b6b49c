"""Helpers shared by tests, bench.py and smoke(): build a lego.yml-shaped pipeline at a small image size,
load seeded synthetic weights.  (Test infrastructure: lives outside the product package; nothing here touches oracle/.)"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch

from tools import synthetic as syn
from yanerf.utils.config import ConfigDict

LEGO_MLP = dict(type="NeRFMLP", n_layers=8, input_skips=[5], n_harmonic_functions_xyz=10,
                harmonic_functions_xyz_append_intput=True, n_hidden_neurons_xyz=256, n_harmonic_functions_dir=4,
                harmonic_functions_dir_append_intput=True, n_hidden_neurons_dir=128, latent_dim=0, input_xyz=True,
                input_dir=True, color_dim=3, nerf_paper_v1=False)


def pipeline_cfg(H: int, W: int, n_rays: int, n_fine: int, noise_std: float, chunk: int,
                 min_depth: float = 2.0, max_depth: float = 6.0, mlp: dict = None, renderer: dict = None) -> ConfigDict:
    """The `pipeline:` block of configs/nerf/lego.yml (lines 45-94) with the size knobs exposed; `renderer` overrides
    entries of the renderer block."""
    cfg = ConfigDict(dict(
        type="NeRFPipeline", chunk_size_grid=chunk, num_passes=2, output_rasterized_mc=True,
        loss_weights={"loss_prev_stage_rgb_mse": 1.0, "loss_rgb_mse": 1.0},
        model=dict(mlp or LEGO_MLP),
        ray_sampler=dict(type="RaySampler", image_height=H, image_width=W, min_depth=min_depth, max_depth=max_depth,
                         n_pts_per_ray_evaluation=64, n_pts_per_ray_training=64,
                         n_rays_per_image_sampled_from_mask=n_rays, scene_extent=0.0,
                         stratified_point_sampling_training=True, stratified_point_sampling_evaluation=False),
        renderer=dict(type="MultipassEmissionAbsorpsionRenderer", append_coarse_samples_to_fine=True,
                      bg_color=[0.0, 0.0, 0.0], blend_output=False, density_noise_std_train=noise_std,
                      n_pts_per_ray_fine_evaluation=n_fine, n_pts_per_ray_fine_training=n_fine,
                      hard_background=False, background_density_bias=1.0e-6),
        feature_extractor=[],
    ))
    cfg["renderer"].update(renderer or {})
    return cfg


def build_pipeline(H, W, n_rays, n_fine, noise_std, chunk, **kw):
    from yanerf.pipelines import PIPELINES

    return PIPELINES.build(pipeline_cfg(H, W, n_rays, n_fine, noise_std, chunk, **kw))


def mlp_param_shapes(mlp: torch.nn.Module) -> Dict[str, tuple]:
    return {k: tuple(v.shape) for k, v in mlp.state_dict().items()}


def load_synth_nets(pipe, seeds: Sequence[int], gain: float) -> List[Dict[str, torch.Tensor]]:
    """Load `synth_mlp_state(seed, gain)` into every implicit function; returns the CPU state dicts
    (the oracle's `nets`)."""
    nets = []
    for fn, seed in zip(pipe.implicit_functions, seeds):
        sd = syn.synth_mlp_state(mlp_param_shapes(fn._fn), seed, gain)
        fn._fn.load_state_dict(sd)
        nets.append(sd)
    return nets
