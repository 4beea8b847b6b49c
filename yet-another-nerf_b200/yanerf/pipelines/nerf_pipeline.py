"""NeRFPipeline: ray_sampler -> feature_extractors -> (chunked) multi-pass renderer -> per-image losses.

API mirror of `yanerf/pipelines/nerf_pipeline.py` (NeRFPipeline 22-324, the chunkify contract 327-426): same
registry name, constructor keywords, keyword-only `forward`, output keys and `(B,)`-shaped losses, so
`scripts/run.py` and `configs/nerf/{lego,fern}.yml` work unchanged.

One deliberate difference, invisible in the results: per-ray work is independent, so with
`coalesce_chunks=True` (default) consecutive `chunk_size_grid` chunks are rendered in one launch sequence
(up to `max_points_per_launch` coarse points) instead of 313 Python-driven iterations; set it to False to
run the reference's exact chunk loop.
"""
from __future__ import annotations

import collections
import dataclasses
import math
from typing import Any, Callable, Dict, List, Optional, Sequence, Union

import torch

from ..pipelines.feature_extractors import FEATURE_EXTRACTORS
from ..pipelines.models import MODELS
from ..pipelines.ray_samplers import RAY_SAMPLERS
from ..pipelines.ray_samplers.utils import RayBundle, RenderSamplingMode
from ..pipelines.renderers import RENDERERS
from ..pipelines.renderers.utils import RendererOutput
from ..pipelines.utils import EvaluationMode, as_mode
from ..utils.logging import get_logger

from .builder import PIPELINES
from .utils import PartialFunctionWrapper, ViewMetrics, sample_grid, scatter_rays_to_image


def chunk_plan(n_rays: int, n_pts_per_ray: int, chunk_size: int):
    """(n_chunks, rays per chunk): `n_chunks = ceil(n_rays * max(P,1) / chunk)`, `ceil(n_rays / n_chunks)`."""
    n_chunks = -(-n_rays * max(n_pts_per_ray, 1) // chunk_size)
    return n_chunks, -(-n_rays // n_chunks)


def _chunk_generator(chunk_size: int, origins, directions, lengths, xys, bg_color=None, *args, **kwargs):
    """Yields `([origins, directions, lengths, xys, bg_color, *args], kwargs)` with every tensor viewed as
    `[B, rays, 1, C]` and sliced along the ray axis."""
    B, *spatial, P = lengths.shape
    n_rays = math.prod(spatial)
    _, per = chunk_plan(n_rays, P, chunk_size)
    flat = lambda t: None if t is None else t.reshape(B, -1, 1, t.shape[-1])
    o, d, z, xy, bg = flat(origins), flat(directions), flat(lengths), flat(xys), flat(bg_color)
    for s in range(0, n_rays, per):
        e = min(s + per, n_rays)
        yield [o[:, s:e], d[:, s:e], z[:, s:e], xy[:, s:e], None if bg is None else bg[:, s:e], *args], kwargs


# ---------------------------------------------------------------------------------------------- ray-slab sharding
# one image over several GPUs.  The flattened H*W ray list is cut into `world` contiguous slabs (a multiple
# of 128 rays, so every slab is a whole number of MLP tiles for any points-per-ray count), every rank renders its slab
# with the replicated weights, then ONE all-gather of `[B, slab, 5]` fp32 per renderer stage (rgb, depth, alpha)
# rebuilds the image on every rank.  No exchange inside a ray.
SLAB_ALIGN = 128


def slab_bounds(n_rays: int, world: int, rank: int, align: int = SLAB_ALIGN):
    """(start, end, rays per slab) of rank's slab; the last slabs may be short or empty."""
    per = -(-n_rays // world)
    per = -(-per // align) * align
    start = min(rank * per, n_rays)
    return start, min(start + per, n_rays), per


def gather_slabs(local: torch.Tensor, n_rays: int, per: int, group=None) -> torch.Tensor:
    """`[B, n_local, C]` (this rank's slab) -> `[B, n_rays, C]` on every rank: ONE `all_gather_into_tensor` into a
    preallocated `[world, B, per, C]` buffer."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local[:, :n_rays]
    world = dist.get_world_size(group)
    B, n_local, C = local.shape
    padded = local if n_local == per else torch.cat((local, local.new_zeros(B, per - n_local, C)), dim=1)
    out = local.new_empty(world, B, per, C)
    dist.all_gather_into_tensor(out.view(-1), padded.contiguous().view(-1), group=group)
    full = out.reshape(world * per, C)[None] if B == 1 else out.permute(1, 0, 2, 3).reshape(B, world * per, C)
    return full[:, :n_rays]


def reduce_sum(t: torch.Tensor, group=None) -> torch.Tensor:
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def _tensor_collator(batch, new_dims) -> torch.Tensor:
    """`[B, rays_i, 1, *rest]` pieces -> `[*new_dims, *rest]`."""
    rest = batch[0].shape[3:]
    joined = batch[0] if len(batch) == 1 else torch.cat(batch, dim=1)
    return joined.reshape(*new_dims, *rest)


def cat_dataclass(batch, tensor_collator: Callable):
    """Field-wise concatenation of a list of (nested) dataclasses; dict fields are concatenated per key."""
    first = batch[0]
    out: Dict[str, Any] = {}
    for f in dataclasses.fields(first):
        v = getattr(first, f.name)
        column = [getattr(e, f.name) for e in batch]
        if v is None:
            out[f.name] = None
        elif torch.is_tensor(v):
            out[f.name] = tensor_collator(column)
        elif dataclasses.is_dataclass(v):
            out[f.name] = cat_dataclass(column, tensor_collator)
        elif isinstance(v, collections.abc.Mapping):
            out[f.name] = {k: (tensor_collator([c[k] for c in column]) if v[k] is not None else None) for k in v}
        else:
            raise ValueError("Unsupported field type for concatenation")
    return type(first)(**out)


def _apply_chunked(func, chunk_generator, tensor_collator):
    return cat_dataclass([func(*a, **kw) for a, kw in chunk_generator], tensor_collator)


@PIPELINES.register_module()
class NeRFPipeline(torch.nn.Module):
    # see module docstring
    coalesce_chunks: bool = True
    max_points_per_launch: int = 64 << 20
    # The reference asserts `xys.max() < image size` on the device values (two device->host syncs per sampled tensor,
    # pipelines/utils.py:283-284).  The pixel grid is produced by this pipeline's own ray sampler, so the same condition is
    # checked on the HOST from the grid and image shapes (`_check_grid_fits`): no sync, and it also fires when a random
    # training pick merely COULD fall outside the image.  True additionally keeps the reference's device-side assertions.
    validate_pixel_grid: bool = False
    # (rank, world, process group) -> full-grid renders are ray-slab sharded over the group (see slab_bounds); None =
    # every rank renders whole images (the reference's DistributedSampler sharding)
    ray_shard: Optional[tuple] = None

    def __init__(
        self,
        ray_sampler,
        model,
        feature_extractor,
        renderer,
        chunk_size_grid: int,
        num_passes: int,
        loss_weights: Dict[str, float] = {"loss_rgb_mse": 1.0, "loss_prev_stage_rgb_mse": 1.0},
        output_rasterized_mc: bool = False,
    ) -> None:
        super().__init__()
        self.logger = get_logger(__name__)
        self.ray_sampler = RAY_SAMPLERS.build(ray_sampler)
        self.render_image_height = ray_sampler["image_height"]
        self.render_image_width = ray_sampler["image_width"]
        self.sampling_mode_training = RenderSamplingMode.MASK_SAMPLE
        self.sampling_mode_evaluation = RenderSamplingMode.FULL_GRID

        if isinstance(model, Sequence) and len(model) != num_passes:
            self.logger.info(f"Rewrite `num_pass` from {num_passes} to {len(model)}.")
            num_passes = len(model)
        self.num_passes = num_passes
        model_cfgs = list(model) if isinstance(model, Sequence) else [model] * num_passes
        self.implicit_functions = torch.nn.ModuleList(PartialFunctionWrapper(MODELS.build(c)) for c in model_cfgs)

        fe_cfgs = list(feature_extractor) if isinstance(feature_extractor, Sequence) else [feature_extractor]
        self.feature_extractors = torch.nn.ModuleList(FEATURE_EXTRACTORS.build(c) for c in fe_cfgs)

        self.renderer = RENDERERS.build(renderer)
        bg_color = renderer["bg_color"] if "bg_color" in renderer else (0.0,)
        if not isinstance(bg_color, torch.Tensor):
            bg_color = torch.tensor(bg_color)
        self.register_buffer("bg_color", bg_color, persistent=False)

        self.chunk_size_grid = chunk_size_grid
        self.output_rasterized_mc = output_rasterized_mc
        self.loss_weights = loss_weights
        self.log_loss_weights()
        self.view_metrics = ViewMetrics()

    def set_device_rng(self, rng) -> None:
        """Hand an `ops.DeviceRng` to every draw site of the training forward (ray sampler, raymarcher, refiners): the
        draws are then generated inside the consuming kernels.  None restores torch's generators (reference behaviour,
        and what the parity tests' draw injection hooks into)."""
        for m in (self.ray_sampler, self.renderer):
            if hasattr(m, "set_device_rng"):
                m.set_device_rng(rng)

    def log_loss_weights(self) -> None:
        rows = "\n".join(f"{k:40s}: {w:1.2e}" for k, w in self.loss_weights.items())
        self.logger.info("-------\nloss_weights:\n" + rows + "\n-------")

    # ------------------------------------------------------------------ forward
    def forward(
        self,
        *,
        poses: torch.Tensor,
        focal_lengths: torch.Tensor,
        image_height: Optional[int] = None,
        image_width: Optional[int] = None,
        min_depth: Optional[float] = None,
        max_depth: Optional[float] = None,
        mask_crop: Optional[torch.Tensor] = None,
        sampling_prob_mask: Optional[torch.Tensor] = None,
        n_rays_per_image: Union[None, int, List[int]] = None,
        bg_image_rgb: Optional[torch.Tensor] = None,
        image_rgb: Optional[torch.Tensor] = None,
        depth_map: Optional[torch.Tensor] = None,
        evaluation_mode: EvaluationMode = EvaluationMode.EVALUATION,
        **kwargs,
    ) -> Dict[str, Any]:
        evaluation_mode = as_mode(evaluation_mode)
        training = evaluation_mode == EvaluationMode.TRAINING
        sampling_mode = RenderSamplingMode(self.sampling_mode_training if training else self.sampling_mode_evaluation)
        masked = sampling_mode == RenderSamplingMode.MASK_SAMPLE
        if sampling_mode == RenderSamplingMode.FULL_GRID and self.ray_shard is not None and depth_map is None:
            return self._forward_sharded(poses, focal_lengths, image_height, image_width, min_depth, max_depth,
                                         bg_image_rgb, image_rgb, evaluation_mode, kwargs)

        ray_bundle: RayBundle = self.ray_sampler(
            poses, focal_lengths, evaluation_mode=evaluation_mode,
            mask=mask_crop if (mask_crop is not None and masked) else None,
            sampling_prob_mask=sampling_prob_mask if training else None,
            n_rays_per_image=n_rays_per_image if training else None,
            image_height=image_height, image_width=image_width, min_depth=min_depth, max_depth=max_depth,
        )
        xys = ray_bundle.xys
        validate = self.validate_pixel_grid
        self._check_grid_fits(image_height, image_width, bg_image_rgb, image_rgb, depth_map)
        bg_color = sample_grid(bg_image_rgb, xys, validate) if bg_image_rgb is not None else None

        extracted = self._extract_features(kwargs)

        for fn in self.implicit_functions:
            fn.bind_args(**extracted)
        deferred = self._defer_weight_checks()
        try:
            rendered: RendererOutput = self._render(
                *ray_bundle, bg_color=bg_color, sampling_mode=sampling_mode,
                implicit_functions=self.implicit_functions, evaluation_mode=evaluation_mode,
            )
        finally:
            for fn in self.implicit_functions:
                fn.unbind_args()
            self._restore_weight_checks(deferred)

        preds = self._get_view_metrics(raymarched=rendered, xys=xys, image_rgb=image_rgb, depth_map=depth_map,
                                       validate_grid=validate)
        blob = {}
        if masked:
            if self.output_rasterized_mc:
                blob = dict(rendered_images=rendered.features, rendered_depths=rendered.depths,
                            rendered_alpha_masks=rendered.alpha_masks)
                blob = self._rasterize_mc_samples(xys, None, image_height, image_width, blob)
        elif sampling_mode == RenderSamplingMode.FULL_GRID:
            blob = dict(rendered_images=rendered.features, rendered_depths=rendered.depths,
                        rendered_alpha_masks=rendered.alpha_masks)
        else:
            raise ValueError(f"Invalid RenderSamplingMode: {sampling_mode}.")
        preds.update(blob)

        objective = self._get_objective(preds)
        if objective is not None:
            preds["objective"] = objective
        self._raise_deferred_weight_checks(deferred)
        return preds

    # ------------------------------------------------------------------ host-side checks without mid-render syncs
    def _check_grid_fits(self, image_height, image_width, *images) -> None:
        custom = image_height is not None and image_width is not None
        gh = image_height if custom else self.render_image_height
        gw = image_width if custom else self.render_image_width
        for t in images:
            if t is not None:
                assert gw <= t.shape[-2], "Invalid ray_sampler.image_width"
                assert gh <= t.shape[-3], "Invalid ray_sampler.image_height"

    def _defer_weight_checks(self):
        """The refiner's "Negative weights provided." check reads a device flag (the reference syncs at the same place,
        renderers/utils.py:123-124).  Inside a pipeline forward the read is moved to the END of the forward: the whole
        render is queued first, the error still surfaces from the same call."""
        refiners = [r for r in getattr(self.renderer, "_refiners", {}).values() if getattr(r, "check_weights", False)]
        for r in refiners:
            r.check_weights = False
            r.last_flag = None
        return refiners

    @staticmethod
    def _restore_weight_checks(refiners) -> None:
        for r in refiners:
            r.check_weights = True

    @staticmethod
    def _raise_deferred_weight_checks(refiners) -> None:
        flags = [r.last_flag for r in refiners if r.last_flag is not None]
        if not flags:
            return
        worst = flags[0] if len(flags) == 1 else torch.stack([f.reshape(()) for f in flags]).max()
        if int(worst.item()) != 0:
            raise ValueError("Negative weights provided.")

    def _extract_features(self, kwargs) -> Dict[str, Any]:
        extracted = collections.defaultdict(list)
        for extractor in self.feature_extractors:
            for k, v in extractor(**kwargs).items():
                extracted[k].append(v)
        for k, vs in extracted.items():
            if isinstance(vs[0], torch.Tensor):
                extracted[k] = torch.stack(vs, dim=1)
            elif len(vs) != 1:
                raise KeyError(f"{k} has multiple {type(vs[0])} values.")
            else:
                extracted[k] = vs[0]
        return extracted

    def _render(self, origins, directions, lengths, xys, *, bg_color, sampling_mode, **kwargs) -> RendererOutput:
        return self._render_local(origins, directions, lengths, xys, bg_color=bg_color, sampling_mode=sampling_mode,
                                  **kwargs)

    def _forward_sharded(self, poses, focal_lengths, image_height, image_width, min_depth, max_depth, bg_image_rgb,
                         image_rgb, evaluation_mode, kwargs) -> Dict[str, Any]:
        """Full-grid forward with the H*W ray list cut into one contiguous slab per rank: this rank
        generates, renders and scores ONLY its slab; then one all-gather of `[slab, 2 x (C + 2)]` (both stages' rgb, depth,
        alpha) rebuilds the images on every rank and one all-reduce of the per-image squared-error sums gives the losses.
        The images are bit-identical to the unsharded render (rays are independent, slabs are whole MLP tiles); the losses
        agree up to the fp32 summation order."""
        from .utils import huber

        rank, world, group = self.ray_shard
        custom = image_height is not None and image_width is not None
        Hh, Ww = (image_height, image_width) if custom else (self.render_image_height, self.render_image_width)
        n_rays = Hh * Ww
        start, end, per = slab_bounds(n_rays, world, rank)
        B = poses.shape[0]
        validate = self.validate_pixel_grid
        self._check_grid_fits(image_height, image_width, bg_image_rgb, image_rgb)
        n_local = end - start
        stages: List[torch.Tensor] = []
        sums: Dict[str, torch.Tensor] = {}
        channels = None
        if n_local > 0:
            bundle: RayBundle = self.ray_sampler(
                poses, focal_lengths, evaluation_mode=evaluation_mode, image_height=image_height, image_width=image_width,
                min_depth=min_depth, max_depth=max_depth, ray_range=(start, end))
            xys = bundle.xys
            bg_color = sample_grid(bg_image_rgb, xys, validate) if bg_image_rgb is not None else None
            extracted = self._extract_features(kwargs)
            for fn in self.implicit_functions:
                fn.bind_args(**extracted)
            deferred = self._defer_weight_checks()
            try:
                out = self._render_local(*bundle, bg_color=bg_color, sampling_mode=RenderSamplingMode.FULL_GRID,
                                         implicit_functions=self.implicit_functions, evaluation_mode=evaluation_mode)
            finally:
                for fn in self.implicit_functions:
                    fn.unbind_args()
                self._restore_weight_checks(deferred)
            local = self._get_view_metrics(raymarched=out, xys=xys, image_rgb=image_rgb, depth_map=None, validate_grid=validate)
            sums = {k: v for k, v in local.items() if k.endswith("rgb_mse")}  # per-slab means
            while out is not None:
                stages += [out.features, out.depths, out.alpha_masks]
                out = out.prev_stage
            channels = sum(t.shape[-1] for t in stages[:3])
            n_stages = len(stages) // 3
            packed = torch.cat(stages, dim=-1).reshape(B, n_local, -1)  # finest stage first; one copy for all of them
        else:  # more ranks than slabs: an empty contribution of the right width
            channels = int(self.bg_color.numel() if self.bg_color.numel() > 1 else 3) + 2
            n_stages = self.num_passes
            packed = poses.new_zeros(B, 0, channels * n_stages)
        full = gather_slabs(packed, n_rays, per, group).reshape(B, Hh, Ww, packed.shape[-1])
        preds: Dict[str, Any] = {}
        if image_rgb is not None:
            keys = ["loss_" + "prev_stage_" * k + "rgb_mse" for k in range(n_stages)]
            # slab mean -> slab sum -> (all-reduce) image sum -> image mean; the hubers of all stages in one go
            acc = torch.stack([sums[k] for k in keys]) * float(n_local) if sums else poses.new_zeros(n_stages, B)
            acc = reduce_sum(acc, group) / float(n_rays)
            hub = huber(acc, scaling=0.03)
            for k, v, h in zip(keys, acc, hub):
                preds[k] = v
                preds[k.replace("rgb_mse", "rgb_huber")] = h
        fine = full[..., :channels]
        preds.update(rendered_images=fine[..., :-2], rendered_depths=fine[..., -2:-1], rendered_alpha_masks=fine[..., -1:])
        objective = self._get_objective(preds)
        if objective is not None:
            preds["objective"] = objective
        if n_local > 0:
            self._raise_deferred_weight_checks(deferred)
        return preds

    def _render_local(self, origins, directions, lengths, xys, *, bg_color, sampling_mode, **kwargs) -> RendererOutput:
        if sampling_mode == RenderSamplingMode.FULL_GRID and self.chunk_size_grid > 0:
            chunk = self.chunk_size_grid
            if self.coalesce_chunks:
                chunk = max(chunk, self.max_points_per_launch)
            return _apply_chunked(
                self.renderer,
                _chunk_generator(chunk, origins, directions, lengths, xys, bg_color, **kwargs),
                lambda pieces: _tensor_collator(pieces, lengths.shape[:-1]),
            )
        return self.renderer(origins=origins, directions=directions, lengths=lengths, xys=xys, bg_color=bg_color, **kwargs)

    # ------------------------------------------------------------------ losses
    def _get_view_metrics(self, raymarched: RendererOutput, xys, image_rgb=None, depth_map=None, keys_prefix: str = "loss_",
                          validate_grid: bool = True):
        metrics = {}
        stage, prefix = raymarched, keys_prefix
        while stage is not None:
            metrics.update(self.view_metrics(
                image_sampling_grid=xys, images_pred=stage.features, images=image_rgb,
                depths_pred=stage.depths, depths=depth_map, keys_prefix=prefix, validate_grid=validate_grid,
            ))
            stage, prefix = stage.prev_stage, prefix + "prev_stage_"
        return metrics

    def _get_objective(self, preds) -> Optional[torch.Tensor]:
        missing = [k for k in self.loss_weights if k not in preds]
        for k in missing:
            self.logger.warning(f"loss name is not found: {k}")
        terms = [preds[k] * float(w) for k, w in self.loss_weights.items() if k in preds and w != 0.0]
        if not terms:
            self.logger.warning("No main objective found.")
            return None
        loss = sum(terms)
        assert torch.is_tensor(loss)
        return loss

    def _rasterize_mc_samples(self, xys, bg_color, image_height, image_width, rendered_dict):
        if image_height is None or image_width is None:
            image_height, image_width = self.render_image_height, self.render_image_width
        vals = list(rendered_dict.values())
        if bg_color is None and xys.is_cuda and 1 <= len(vals) <= 3:
            # one zero fill + one launch for rgb / depth / alpha together (`yn_scatter_rays`)
            from .. import ops

            with torch.no_grad():
                B = xys.shape[0]
                outs = ops.scatter_rays([v.detach() for v in vals], xys.reshape(B, -1, 2), image_height, image_width)
            return dict(zip(rendered_dict.keys(), outs))
        return {k: scatter_rays_to_image(v, xys, image_height, image_width, bg_color) for k, v in rendered_dict.items()}
