"""Thin re-host of the reference's runner loop (`yanerf/runners/apis.py:30-236`, `runners/utils.py:257-283`) on
top of `FusedTrainer`: same call structure (`inference` -> `model(**data, evaluation_mode=...)`), same stats
(`create_stats`: mean of every loss_/objective key, PSNR from every "mse" key), same eval sharding (one image per
rank, all-gather of the `(B,)` losses, truncation to the dataset length).  Data loading, PNG dumps, hooks and
checkpoint rotation stay with the caller (`scripts/run.py`), as in the reference."""
from __future__ import annotations

import math
from typing import Any, Dict, Iterable, Optional

import torch
import torch.distributed as dist

from ..pipelines.utils import EvaluationMode

from .engine import FusedTrainer, reference_lr, scaled_runner_config


def inference(model: torch.nn.Module, data: Dict[str, Any], evaluation_mode: EvaluationMode, compute_metrics: bool = True):
    return model(**data, evaluation_mode=evaluation_mode)


def create_stats(preds: Dict[str, Any]) -> Dict[str, float]:
    stats = {}
    for k, v in preds.items():
        if torch.is_tensor(v) and (k.startswith("loss_") or k.startswith("objective")):
            stats[k] = float(v.detach().float().mean())
    for k in [k for k in stats if "mse" in k]:
        stats[k.replace("mse", "psnr")] = -10.0 * math.log10(max(stats[k], 1e-12))
    return stats


@torch.no_grad()
def concat_all_gather(t: torch.Tensor) -> torch.Tensor:
    """all_gather along dim 0 (identity without an initialised process group)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, t.contiguous())
    return torch.cat(parts, dim=0)


def enable_ray_sharding(model: torch.nn.Module, group=None) -> bool:
    """SURVEY 8(e), render: from now on every full-grid render of `model` is split into contiguous ray slabs over the
    ranks of `group` and re-assembled with one all-gather per renderer stage (`NeRFPipeline.ray_shard`).  All ranks must
    then call the model with the SAME image.  Returns False (and changes nothing) without a multi-rank process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return False
    model.ray_shard = (dist.get_rank(group), dist.get_world_size(group), group)
    return True


def train_one_epoch(trainer: FusedTrainer, batches: Iterable[Dict[str, Any]], config: Dict[str, Any], epoch: int = 0,
                    iters_per_epoch: Optional[int] = None) -> Dict[str, float]:
    """One pass over `batches` (dicts of device tensors keyed like the dataset NamedTuples).

    `config` is the reference's runner config (configs/nerf/lego.yml:12-33): `init_lr`, `min_lr`, `lr_decay_type`,
    `lr_decay_rate`, `lr_decay_iters`, `num_iters`, `warmup_steps`, `warmup_lr`, `linear_scale`.  The learning rate of
    every iteration is the reference's (`reference_lr`: decay schedule, then the warm-up override while
    `iter <= warmup_steps`, runners/apis.py:77-79), with `init_lr` / `min_lr` scaled by the world size when
    `linear_scale` is set and a process group is initialised (scripts/run.py:152-156)."""
    missing = [k for k in ("init_lr", "min_lr") if k not in config]
    if missing:
        raise KeyError(f"runner config lacks {missing} (reference keys: init_lr, min_lr, lr_decay_type, lr_decay_rate, "
                       "lr_decay_iters, num_iters, warmup_steps, warmup_lr, linear_scale)")
    distributed = dist.is_available() and dist.is_initialized()
    cfg = scaled_runner_config(dict(config), trainer.world, distributed)
    trainer.init_lr = cfg["init_lr"]  # written into checkpoints (the reference's param groups carry it)
    it = epoch * (iters_per_epoch or 0)
    trainer.pipeline.train()
    preds: Dict[str, Any] = {}
    for data in batches:
        preds = trainer.train_step(data, lr=reference_lr(it, **cfg))
        it += 1
    return create_stats(preds)


@torch.no_grad()
def eval_one_epoch(model: torch.nn.Module, batches: Iterable[Dict[str, Any]], dataset_len: Optional[int] = None) -> Dict[str, float]:
    model.eval()
    gathered: Dict[str, list] = {}
    for data in batches:
        preds = inference(model, data, EvaluationMode.EVALUATION)
        for k, v in preds.items():
            if torch.is_tensor(v) and (k.startswith("loss_") or k.startswith("objective")):
                gathered.setdefault(k, []).append(concat_all_gather(v))
    out = {}
    for k, vs in gathered.items():
        allv = torch.cat(vs, dim=0)
        if dataset_len is not None:
            allv = allv[:dataset_len]
        out[k] = allv
    return create_stats(out)
