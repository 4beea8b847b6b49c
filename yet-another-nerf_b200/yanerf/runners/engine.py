"""Training step engine: the device side of `runners/apis.py:train_one_epoch` (53-118) for one iteration.

`scripts/run.py` builds `torch.optim.Adam` over 48 tensors and wraps the pipeline in DDP (run.py:152-166); here the
same arithmetic runs on flat buffers:
  * every parameter (and its .grad) is a view into ONE contiguous fp32 buffer;
  * the DDP gradient all-reduce (mean) is a single NCCL all-reduce of that flat gradient buffer;
  * Adam is one `yn_adam_step` launch over the flat buffers (the 1/world_size of the mean folded in as grad_scale).
The forward/backward itself is `NeRFPipeline.forward` + autograd, i.e. the kernels of ops.py.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional

import torch
import torch.distributed as dist

from yanerf import ops
from yanerf.pipelines.utils import EvaluationMode


def exponential_lr(it: int, lr: float, min_lr: float, num_iters: int, warmup_iters: int = 0, warmup_lr: float = 0.0) -> float:
    """Learning-rate schedule of the reference runner (runners/utils.py:65-109): linear warm-up from
    `warmup_lr`, then exponential decay lr -> min_lr over `num_iters`."""
    if warmup_iters > 0 and it < warmup_iters:
        return warmup_lr + (lr - warmup_lr) * it / warmup_iters
    t = min(max(it, 0), num_iters) / max(num_iters, 1)
    return lr * (min_lr / lr) ** t


class FusedTrainer:
    """Owns flat parameter / gradient / Adam-moment buffers of a pipeline and performs one optimisation step."""

    def __init__(self, pipeline: torch.nn.Module, lr: float = 5e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 process_group: Optional[Any] = None, use_cuda_graph: bool = False, graph_warmup_steps: int = 3) -> None:
        """use_cuda_graph: after `graph_warmup_steps` eager steps the whole iteration (pixel pick, weight re-pack,
        forward, backward, Adam) is captured once and replayed; step counter, learning rate and the pixel seed live in
        device memory so every replay sees fresh values.  With several ranks the NCCL all-reduce and Adam follow the
        replay as ordinary launches."""
        self.pipeline = pipeline
        self.use_cuda_graph = use_cuda_graph
        self._graph_warmup = graph_warmup_steps
        self._graph = None
        self._side_stream = None
        self._warm_steps = 0
        self._static_batch: Dict[str, Any] = {}
        self._static_preds: Dict[str, Any] = {}
        self.lr, self.betas, self.eps = lr, betas, eps
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
        self._lr_host = torch.zeros(1).pin_memory() if dev.type == "cuda" else torch.zeros(1)
        for m in pipeline.modules():  # O(n) graph-friendly pixel pick instead of torch.multinomial over H*W
            if hasattr(m, "fused_pixel_sampler"):
                m.fused_pixel_sampler = True
        self._flag_host, self._flag_event, self._flag_pending = None, None, False
        # no device->host sync inside the step: pixel-grid range checks move to the host (shapes) and the refiner's
        # "Negative weights provided." flag is read once, after the whole step has been queued
        pipeline.validate_pixel_grid = False
        self._refiners = [m for m in pipeline.modules() if hasattr(m, "check_weights")]
        for r in getattr(getattr(pipeline, "renderer", None), "_refiners", {}).values():
            if r not in self._refiners:
                self._refiners.append(r)
        for r in self._refiners:
            r.check_weights = False
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

    def _eager_step(self, batch: Dict[str, Any], lr: Optional[float]) -> Dict[str, torch.Tensor]:
        self.zero_grad()
        preds = self.pipeline(**batch, evaluation_mode=EvaluationMode.TRAINING)
        if "objective" not in preds:
            raise KeyError("objective")  # runners/apis.py:90-91
        preds["objective"].mean().backward()
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
        self._lr_host[0] = self.lr if lr is None else lr
        if self._graph is None:
            self._capture(batch)
        for k, v in batch.items():
            if torch.is_tensor(v):
                self._static_batch[k].copy_(v, non_blocking=True)
        self._state[1:2].copy_(self._lr_host, non_blocking=True)
        self._graph.replay()  # (the captured forward starts by re-packing the weight images from the flat buffer)
        if self.world > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
            self._graph_optimizer()
        self.step_count += 1
        self._invalidate()    # other weight images (e.g. the fp16 one used by evaluation) are stale now
        self._deferred_checks()
        return self._static_preds

    def _capture(self, batch: Dict[str, Any]) -> None:
        self._static_batch = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}
        self._state[0] = float(self.step_count)
        self._state[1:2].copy_(self._lr_host, non_blocking=True)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph, stream=self._side_stream):
            self.flat_grad.zero_()
            preds = self.pipeline(**self._static_batch, evaluation_mode=EvaluationMode.TRAINING)
            if "objective" not in preds:
                raise KeyError("objective")
            preds["objective"].mean().backward()
            if self.world == 1:
                self._graph_optimizer()
            self._static_preds = {k: v.detach() if torch.is_tensor(v) else v for k, v in preds.items()}

    def _graph_optimizer(self) -> None:
        """Adam with step count / learning rate read from device memory.  Single GPU: part of the captured graph.
        Multi-GPU: launched after the replay, behind the NCCL all-reduce (NCCL stays outside the capture: its
        watchdog and the capture do not mix reliably)."""
        self._state[0:1] += 1.0
        ops.adam_step_dev(self.flat, self.flat_grad, self.exp_avg, self.exp_avg_sq, self._state, self.betas[0],
                          self.betas[1], self.eps, grad_scale=1.0 / self.world)

    def _deferred_checks(self) -> None:
        """The refiner's device flag of step k is copied to pinned memory asynchronously and examined at the end of
        step k+1 (or by `finish()`), so the host never waits for the GPU inside a step."""
        self._raise_if_flagged(wait=False)
        flags = [r.last_flag for r in self._refiners if r.last_flag is not None]
        if flags:
            if self._flag_host is None:
                self._flag_host = torch.zeros(1, dtype=torch.int32).pin_memory()
                self._flag_event = torch.cuda.Event()
            self._flag_host.copy_(torch.stack([f.reshape(()) for f in flags]).max().reshape(1), non_blocking=True)
            self._flag_event.record()
            self._flag_pending = True

    def _raise_if_flagged(self, wait: bool) -> None:
        if not self._flag_pending:
            return
        if wait:
            self._flag_event.synchronize()
        elif not self._flag_event.query():
            return
        self._flag_pending = False
        if int(self._flag_host.item()) != 0:
            raise ValueError("Negative weights provided.")  # renderers/utils.py:123-124

    def finish(self) -> None:
        """Wait for outstanding work and surface any deferred error."""
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
    # reference checkpoint resumes here and a checkpoint written here resumes in the reference.
    def optimizer_state_dict(self) -> Dict[str, Any]:
        state, off = {}, 0
        for i, p in enumerate(self.params):
            k = p.numel()
            state[i] = {"step": torch.tensor(float(self.step_count)),
                        "exp_avg": self.exp_avg[off:off + k].view_as(p), "exp_avg_sq": self.exp_avg_sq[off:off + k].view_as(p)}
            off += k
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": 0, "amsgrad": False,
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
        order = [i for g in groups for i in g["params"]]
        if len(order) != len(self.params):
            raise ValueError(f"optimizer state for {len(order)} parameters, the pipeline has {len(self.params)}")
        if groups[0].get("weight_decay", 0) or groups[0].get("amsgrad", False):
            raise NotImplementedError("weight_decay / amsgrad checkpoints (the reference trains with plain Adam, run.py:159)")
        self.lr = float(groups[0].get("lr", self.lr))
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
