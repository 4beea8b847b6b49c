"""Training step engine: the device side of `runners/apis.py:train_one_epoch` (53-118) for one iteration.

`scripts/run.py` builds `torch.optim.Adam` over 48 tensors and wraps the pipeline in DDP (run.py:152-166); here the
same arithmetic runs on flat buffers:
  * every parameter (and its .grad) is a view into ONE contiguous fp32 buffer;
  * the DDP gradient all-reduce (mean) is a single NCCL all-reduce of that flat gradient buffer;
  * Adam is one `yn_adam_step` launch over the flat buffers (the 1/world_size of the mean folded in as grad_scale).
The forward/backward itself is `NeRFPipeline.forward` + autograd, i.e. the kernels of ops.py.
"""
from __future__ import annotations

import math
from typing import Any, Callable, Dict, List, Optional

import torch
import torch.distributed as dist

from .. import ops
from ..pipelines.utils import EvaluationMode


def reference_lr(it: int, *, init_lr: float, min_lr: float, lr_decay_type: str = "exponential", lr_decay_rate: float = 0.1,
                 lr_decay_iters: int = 250000, num_iters: int = 200000, warmup_steps: int = 0, warmup_lr: float = 0.0,
                 **_unused) -> float:
    """Learning rate of iteration `it`, exactly as the reference runner sets it before every step
    (runners/apis.py:77-79): the decay schedule first --
      "exponential": `max(min_lr, init_lr * lr_decay_rate ** (it / lr_decay_iters))`   (runners/utils.py:82-86)
      "cosine":      `(init_lr - min_lr) * 0.5 * (1 + cos(pi * (it / lr_decay_iters) / num_iters)) + min_lr`   (73-79)
    -- then, while `it <= warmup_steps`, overwritten by the linear warm-up
      `min(init_lr, warmup_lr + (init_lr - warmup_lr) * it / warmup_steps)`   (65-70).
    Keyword names are the runner-config keys of configs/nerf/lego.yml:12-33."""
    if lr_decay_type == "exponential":
        lr = max(min_lr, init_lr * (lr_decay_rate ** (it / lr_decay_iters)))
    elif lr_decay_type == "cosine":
        lr = (init_lr - min_lr) * 0.5 * (1.0 + math.cos(math.pi * (it / lr_decay_iters) / num_iters)) + min_lr
    else:
        raise ValueError(f"lr_decay_type {lr_decay_type!r}")  # runners/utils.py:106-107
    if warmup_steps > 0 and it <= warmup_steps:
        lr = min(init_lr, warmup_lr + (init_lr - warmup_lr) * it / warmup_steps)
    return lr


def scaled_runner_config(config: Dict[str, Any], world_size: int, distributed: bool) -> Dict[str, Any]:
    """`scripts/run.py:152-156`: with an initialised process group and `linear_scale` set, `init_lr` and `min_lr` are
    multiplied by the world size (the warm-up start `warmup_lr` is not)."""
    out = dict(config)
    if distributed and out.get("linear_scale", False):
        out["init_lr"] = out["init_lr"] * world_size
        out["min_lr"] = out["min_lr"] * world_size
    return out


class FusedTrainer:
    """Owns flat parameter / gradient / Adam-moment buffers of a pipeline and performs one optimisation step."""

    def __init__(self, pipeline: torch.nn.Module, lr: float = 5e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 process_group: Optional[Any] = None, use_cuda_graph: bool = False, graph_warmup_steps: int = 3,
                 device_rng: bool = True) -> None:
        """use_cuda_graph: after `graph_warmup_steps` eager steps the whole iteration (pixel pick, weight re-pack,
        forward, backward, Adam) is captured once and replayed; step counter, learning rate and the pixel seed live in
        device memory so every replay sees fresh values.  With several ranks the NCCL all-reduce and Adam follow the
        replay as ordinary launches.
        device_rng: the draws of the training forward (pixel pick, stratified jitter, density noise, inverse-CDF uniforms)
        are generated inside the consuming kernels from one device-resident Philox state (`ops.DeviceRng`, seeded from
        torch's CPU generator) instead of `torch.multinomial / rand / randn` tensors; False keeps torch's generators."""
        self.pipeline = pipeline
        self.use_cuda_graph = use_cuda_graph
        self._graph_warmup = graph_warmup_steps
        self._graph = None
        self._side_stream = None
        self._warm_steps = 0
        self._static_batch: Dict[str, Any] = {}
        self._static_preds: Dict[str, Any] = {}
        self.lr, self.betas, self.eps = lr, betas, eps
        self.init_lr = lr  # the reference's param groups carry `init_lr` (runners/utils.py:151,181); kept in checkpoints
        if use_cuda_graph and graph_warmup_steps < 1:
            raise ValueError("graph_warmup_steps must be >= 1: the first eager step fills the host-side caches "
                             "(linspace rows, kernel attributes) that must not be part of the capture")
        self.group = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.params: List[torch.nn.Parameter] = [p for p in pipeline.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("pipeline has no trainable parameters")
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.empty(n, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(n, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        off = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat[off:off + k].copy_(p.reshape(-1))
                p.data = self.flat[off:off + k].view_as(p)
                p.grad = self.flat_grad[off:off + k].view_as(p)
                off += k
        self.step_count = 0
        self._state = torch.zeros(2, device=dev)  # [step, lr] read by yn_adam_step_dev
        # pinned staging ring for the per-step learning rate: slot k is rewritten only after the copy that read it has
        # completed (its event), so an in-flight non_blocking H2D copy never sees the next step's value
        self._lr_ring = [[torch.zeros(1).pin_memory() if dev.type == "cuda" else torch.zeros(1), None] for _ in range(4)]
        self._lr_slot = 0
        for m in pipeline.modules():  # O(n) graph-friendly pixel pick instead of torch.multinomial over H*W
            if hasattr(m, "fused_pixel_sampler"):
                m.fused_pixel_sampler = True
        self.rng = ops.DeviceRng(dev) if (device_rng and dev.type == "cuda" and hasattr(pipeline, "set_device_rng")) else None
        if self.rng is not None:
            pipeline.set_device_rng(self.rng)
        # no device->host sync inside the step: pixel-grid range checks move to the host (shapes) and the refiner's
        # "Negative weights provided." flag is ACCUMULATED on the device (flag |= this step's flags) and copied to pinned
        # memory after every step; the host looks at copies whose event has completed, so a flag raised by step k can
        # be reported late but never lost.  The switches are flipped only while a training step runs.
        self._flag_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self._flag_slots = ([[torch.zeros(1, dtype=torch.int32).pin_memory(), torch.cuda.Event()] for _ in range(8)]
                            if dev.type == "cuda" else [])
        self._flag_ring: List[int] = []  # indices of slots whose copy is in flight, oldest first
        self._refiners = [m for m in pipeline.modules() if hasattr(m, "check_weights")]
        for r in getattr(getattr(pipeline, "renderer", None), "_refiners", {}).values():
            if r not in self._refiners:
                self._refiners.append(r)
        self._mlps = [m for m in pipeline.modules() if hasattr(m, "invalidate_packed_weights")]
        # hand every NeRFMLP its slice of the flat buffers: one autograd leaf per network, gradients accumulated by
        # the kernels directly into flat_grad
        offsets = {}
        off = 0
        for p in self.params:
            offsets[id(p)] = off
            off += p.numel()
        for m in self._mlps:
            ps = m.ordered_parameters()
            o0, k = offsets[id(ps[0])], sum(p.numel() for p in ps)
            m.use_flat_parameters(self.flat[o0:o0 + k].detach().requires_grad_(True), self.flat_grad[o0:o0 + k])
        self._broadcast_parameters()

    def _broadcast_parameters(self) -> None:
        if self.world > 1:
            dist.broadcast(self.flat, src=0, group=self.group)
            self._invalidate()

    def _invalidate(self) -> None:
        for m in self._mlps:
            m.invalidate_packed_weights()

    def zero_grad(self) -> None:
        self.flat_grad.zero_()

    def eager_step(self, batch: Dict[str, Any], lr: Optional[float] = None) -> Dict[str, torch.Tensor]:
        """One step outside the CUDA graph (same kernels, launched one by one)."""
        preds = self._eager_step(batch, lr)
        self._deferred_checks()
        return preds

    class _NoHostSync:
        """While a training step is being queued the pipeline must not read device values on the host: the pixel-grid
        range assert and the refiner's `weights.min() <= 0` check are switched off (the latter is replaced by the
        device flag) and restored afterwards, so evaluation keeps the reference's immediate errors."""

        def __init__(self, trainer: "FusedTrainer") -> None:
            self.t = trainer

        def __enter__(self):
            t = self.t
            self.saved = (getattr(t.pipeline, "validate_pixel_grid", True), [r.check_weights for r in t._refiners])
            t.pipeline.validate_pixel_grid = False
            for r in t._refiners:
                r.check_weights = False

        def __exit__(self, *exc):
            t = self.t
            t.pipeline.validate_pixel_grid = self.saved[0]
            for r, v in zip(t._refiners, self.saved[1]):
                r.check_weights = v

    def _forward_backward(self, batch: Dict[str, Any]) -> Dict[str, torch.Tensor]:
        with FusedTrainer._NoHostSync(self):
            preds = self.pipeline(**batch, evaluation_mode=EvaluationMode.TRAINING)
            if "objective" not in preds:
                raise KeyError("In train mode, but no loss (`objective`) is found.")  # runners/apis.py:90-91
            preds["objective"].mean().backward()
            for r in self._refiners:  # accumulate this step's "weight + 1e-5 <= 0 seen" flags on the device
                if r.last_flag is not None:
                    self._flag_dev.bitwise_or_(r.last_flag.reshape(1).to(torch.int32))
        return preds

    def _eager_step(self, batch: Dict[str, Any], lr: Optional[float]) -> Dict[str, torch.Tensor]:
        self.zero_grad()
        if self.rng is not None:
            ops.step_begin(self.rng, None)
        preds = self._forward_backward(batch)
        self.optimizer_step(lr)
        return preds

    def train_step(self, batch: Dict[str, Any], lr: Optional[float] = None) -> Dict[str, torch.Tensor]:
        """forward (TRAINING) -> objective.mean().backward() -> all-reduce -> Adam.  Returns the preds dict (with
        `use_cuda_graph` the same static tensors every step: read them before the next call)."""
        if not self.use_cuda_graph or self.flat.device.type != "cuda":
            preds = self._eager_step(batch, lr)
            self._deferred_checks()
            return preds
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream()
        if self._graph is None and self._warm_steps < self._graph_warmup:
            # warm-up on the capture stream (autograd remembers the stream a node was first run on): fills the
            # linspace / allocator caches and configures the kernels
            self._warm_steps += 1
            self._side_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._side_stream):
                preds = self._eager_step(batch, lr)
            torch.cuda.current_stream().wait_stream(self._side_stream)
            self._deferred_checks()
            return preds
        reason = self._graph_unsupported(batch)
        if reason is not None:
            raise NotImplementedError(f"use_cuda_graph: {reason}; construct FusedTrainer(use_cuda_graph=False) for this data")
        if self._graph is None:
            self._capture(batch)
        self._check_static(batch)
        for k, v in batch.items():
            if torch.is_tensor(v):
                self._static_batch[k].copy_(v, non_blocking=True)
        self._stage_lr(self.lr if lr is None else lr)
        self._graph.replay()  # (the captured forward starts by re-packing the weight images from the flat buffer)
        if self.world > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
            self._graph_optimizer()
        self.step_count += 1
        self._invalidate()    # other weight images (e.g. the fp16 one used by evaluation) are stale now
        self._deferred_checks()
        return self._static_preds

    def _stage_lr(self, lr: float) -> None:
        host, ev = self._lr_ring[self._lr_slot]
        if ev is not None:
            ev.synchronize()  # the copy that last read this slot (4 steps ago) has long completed
        host[0] = lr
        self._state[1:2].copy_(host, non_blocking=True)
        if self.flat.device.type == "cuda":
            if ev is None:
                ev = self._lr_ring[self._lr_slot][1] = torch.cuda.Event()
            ev.record()
        self._lr_slot = (self._lr_slot + 1) % len(self._lr_ring)

    def _graph_unsupported(self, batch: Dict[str, Any]) -> Optional[str]:
        """Inputs whose handling needs a device->host read inside the step (`.item()` in the ray sampler:
        ray_sampler.py:179, 280-283) cannot be captured: per-image depth bounds given as tensors (LLFF), sampling
        masks, `scene_extent > 0`."""
        for k in ("min_depth", "max_depth", "mask", "sampling_prob_mask", "fg_probability"):
            if torch.is_tensor(batch.get(k)):
                return f"batch field `{k}` is a tensor that the ray sampler reduces on the host"
        rs = getattr(self.pipeline, "ray_sampler", None)
        if rs is not None and float(getattr(rs, "scene_extent", 0.0) or 0.0) > 0:
            return "ray_sampler.scene_extent > 0 derives the depth range from the camera position on the host"
        return None

    def _check_static(self, batch: Dict[str, Any]) -> None:
        """Every replay must see the batch structure that was captured: same keys, tensor shapes / dtypes, and equal
        non-tensor values (those were baked into the graph)."""
        if set(batch) != set(self._static_batch):
            raise ValueError(f"batch keys changed after graph capture: {sorted(batch)} vs {sorted(self._static_batch)}")
        for k, v in batch.items():
            ref = self._static_batch[k]
            if torch.is_tensor(v):
                if not torch.is_tensor(ref) or v.shape != ref.shape or v.dtype != ref.dtype:
                    raise ValueError(f"batch field `{k}`: {tuple(v.shape)} {v.dtype} does not match the captured "
                                     f"{tuple(ref.shape) if torch.is_tensor(ref) else type(ref).__name__}")
            elif v != ref:
                raise ValueError(f"batch field `{k}` = {v!r} differs from the value captured in the graph ({ref!r})")

    def _capture(self, batch: Dict[str, Any]) -> None:
        self._static_batch = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}
        self._state[0] = float(self.step_count)
        # whatever weight image an earlier (e.g. evaluation) forward left behind, the captured forward must START with
        # the re-pack from the flat buffer: drop every cached image so that plan_for() packs inside the capture
        self._invalidate()
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph, stream=self._side_stream):
            self.flat_grad.zero_()
            ops.step_begin(self.rng, self._state)  # snapshot of the draw step + Adam's device-side step counter
            preds = self._forward_backward(self._static_batch)
            if self.world == 1:
                self._graph_optimizer()
            self._static_preds = {k: v.detach() if torch.is_tensor(v) else v for k, v in preds.items()}

    def _graph_optimizer(self) -> None:
        """Adam with step count (bumped by `yn_step_begin` at the start of the captured step) / learning rate read from
        device memory.  Single GPU: part of the captured graph.
        Multi-GPU: launched after the replay, behind the NCCL all-reduce (NCCL stays outside the capture: its
        watchdog and the capture do not mix reliably)."""
        ops.adam_step_dev(self.flat, self.flat_grad, self.exp_avg, self.exp_avg_sq, self._state, self.betas[0],
                          self.betas[1], self.eps, grad_scale=1.0 / self.world)

    def _deferred_checks(self) -> None:
        """Queue an asynchronous copy of the accumulated device flag into a fresh pinned slot and examine the slots
        whose copies have completed.  The host never waits for the GPU inside a step; `finish()` waits."""
        self._raise_if_flagged(wait=False)
        if self.flat.device.type != "cuda":
            if int(self._flag_dev.item()) != 0:
                self._flag_dev.zero_()
                raise ValueError("Negative weights provided.")
            return
        if len(self._flag_ring) == len(self._flag_slots):  # every slot in flight: wait for the oldest copy
            self._raise_if_flagged(wait=True, only_oldest=True)
        i = next(k for k in range(len(self._flag_slots)) if k not in self._flag_ring)
        host, ev = self._flag_slots[i]
        host.copy_(self._flag_dev, non_blocking=True)
        ev.record()
        self._flag_ring.append(i)

    def _raise_if_flagged(self, wait: bool, only_oldest: bool = False) -> None:
        while self._flag_ring:
            host, ev = self._flag_slots[self._flag_ring[0]]
            if wait:
                ev.synchronize()
            elif not ev.query():
                return
            self._flag_ring.pop(0)
            if int(host.item()) != 0:
                torch.cuda.current_stream().synchronize()  # outstanding copies must not land in reused slots
                self._flag_ring.clear()
                self._flag_dev.zero_()  # reported: start over
                raise ValueError("Negative weights provided.")  # renderers/utils.py:123-124
            if only_oldest:
                return

    def finish(self) -> None:
        """Wait for outstanding work and surface any deferred error."""
        if self.flat.device.type == "cuda":
            torch.cuda.current_stream().synchronize()
        self._deferred_checks()
        self._raise_if_flagged(wait=True)

    def optimizer_step(self, lr: Optional[float] = None) -> None:
        if self.world > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
        self.step_count += 1
        ops.adam_step(self.flat, self.flat_grad, self.exp_avg, self.exp_avg_sq, self.lr if lr is None else lr,
                      self.step_count, self.betas[0], self.betas[1], self.eps, grad_scale=1.0 / self.world)
        if self._graph is not None:
            self._state[0:1].fill_(float(self.step_count))  # keep the graph's device-side step counter in sync
        self._invalidate()

    # ------------------------------------------------------------------ checkpoints (scripts/run.py:168-178, 416-422)
    # Wire format of the reference: {"model": pipeline.state_dict(), "optimizer": torch.optim.Adam.state_dict(),
    # "epoch": e}.  The optimizer state is indexed by parameter position in `model.parameters()` order, which is the
    # order of `self.params`; the moments are views into the flat buffers on the way out and copied in on the way in, so a
    # reference checkpoint resumes here and a checkpoint written here resumes in the reference: the single param group
    # carries `init_lr` like the reference's (`create_param_groups`, runners/utils.py:142-186), which every reference
    # scheduler reads after `optimizer.load_state_dict` has replaced the groups with the saved ones.
    def optimizer_state_dict(self) -> Dict[str, Any]:
        state, off = {}, 0
        for i, p in enumerate(self.params):
            k = p.numel()
            state[i] = {"step": torch.tensor(float(self.step_count)),
                        "exp_avg": self.exp_avg[off:off + k].view_as(p), "exp_avg_sq": self.exp_avg_sq[off:off + k].view_as(p)}
            off += k
        group = {"lr": self.lr, "init_lr": self.init_lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": False, "params": list(range(len(self.params)))}
        return {"state": state if self.step_count > 0 else {}, "param_groups": [group]}

    def load_optimizer_state_dict(self, opt: Dict[str, Any]) -> None:
        if "param_groups" not in opt:  # round-1 private layout {"exp_avg", "exp_avg_sq", "step"}
            self.exp_avg.copy_(opt["exp_avg"])
            self.exp_avg_sq.copy_(opt["exp_avg_sq"])
            self.step_count = int(opt["step"])
            return
        groups = opt["param_groups"]
        if len(groups) > 1:
            # runners/utils.py:153-186 (`lr_param_groups`): per-prefix learning rates.  One flat Adam launch has one lr.
            raise NotImplementedError(f"checkpoint with {len(groups)} optimizer param groups (`lr_param_groups`): "
                                      "FusedTrainer keeps a single learning rate for all parameters")
        order = [i for g in groups for i in g["params"]]
        if len(order) != len(self.params):
            raise ValueError(f"optimizer state for {len(order)} parameters, the pipeline has {len(self.params)}")
        if groups[0].get("weight_decay", 0) or groups[0].get("amsgrad", False):
            raise NotImplementedError("weight_decay / amsgrad checkpoints (the reference trains with plain Adam, run.py:159)")
        self.lr = float(groups[0].get("lr", self.lr))
        self.init_lr = float(groups[0].get("init_lr", self.lr))
        self.betas = tuple(groups[0].get("betas", self.betas))
        self.eps = float(groups[0].get("eps", self.eps))
        state, off, steps = opt.get("state", {}), 0, set()
        for pos, p in zip(order, self.params):
            k = p.numel()
            st = state.get(pos, state.get(str(pos)))
            if st is None:  # a parameter that never received a gradient
                self.exp_avg[off:off + k].zero_()
                self.exp_avg_sq[off:off + k].zero_()
            else:
                if tuple(st["exp_avg"].shape) != tuple(p.shape):
                    raise ValueError(f"optimizer state {pos}: shape {tuple(st['exp_avg'].shape)} vs parameter {tuple(p.shape)}")
                self.exp_avg[off:off + k].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
                steps.add(int(float(st["step"])))
            off += k
        if len(steps) > 1:
            raise NotImplementedError(f"per-parameter step counts differ: {sorted(steps)}")
        self.step_count = steps.pop() if steps else 0

    def state_dict(self, epoch: Optional[int] = None) -> Dict[str, Any]:
        out = {"model": self.pipeline.state_dict(), "optimizer": self.optimizer_state_dict()}
        if epoch is not None:
            out["epoch"] = epoch
        return out

    def load_state_dict(self, state: Dict[str, Any]) -> int:
        """Loads a checkpoint in the reference's layout; returns the epoch to resume from (run.py:176)."""
        self.pipeline.load_state_dict(state["model"])  # copies into the flat-buffer views
        if state.get("optimizer"):
            self.load_optimizer_state_dict(state["optimizer"])
        self._state[0:1].fill_(float(self.step_count))
        self._invalidate()
        return int(state.get("epoch", -1)) + 1
