"""`bench.py --workload train`: the lego.yml training step (BASELINE.json configs[2]): 4096 rays per GPU,
64 coarse + 192 fine points per ray, two 8x256 MLPs forward + backward, Adam, gradient all-reduce for N > 1."""
from __future__ import annotations

import json
import os

import torch
import torch.distributed as dist


def run_train_bench(args, rank: int, world: int, dev) -> None:
    import bench as B
    from yanerf import ops
    from yanerf.runners.engine import FusedTrainer

    n_rays = 4096
    peaks = B.load_peaks()
    pipe, _ = B.build_lego_pipeline(dev, n_rays=n_rays)
    use_graph = os.environ.get("YANERF_TRAIN_GRAPH", "1") != "0"
    trainer = FusedTrainer(pipe, lr=5e-4 * world, use_cuda_graph=use_graph)  # linear LR scaling, scripts/run.py:152-156
    poses, focal, image = B.synthetic_inputs(rank)
    batch_d = dict(poses=poses.to(dev), focal_lengths=focal.to(dev), image_rgb=image.to(dev))
    host = dict(poses=poses.pin_memory(), focal_lengths=focal.pin_memory(), image_rgb=image.pin_memory())
    loss_h = torch.empty(1).pin_memory()
    h2d = sum(t.numel() * 4 for t in host.values())

    def step_resident():
        return trainer.train_step(batch_d)

    def step_e2e():
        b = host if use_graph else {k: v.to(dev, non_blocking=True) for k, v in host.items()}  # graph: H2D into static inputs
        preds = trainer.train_step(b)
        loss_h.copy_(preds["objective"], non_blocking=True)
        return preds

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(5, args.warmup)):  # includes the 3 eager steps before the graph is captured
        step_resident()
    torch.cuda.synchronize()
    sampler = B.ClockSampler(dev.index or 0)
    if rank == 0:
        sampler.start()
    total_ms = timed(step_resident, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(2):
        step_e2e()
    e2e_ms = timed(step_e2e, args.steps)
    # per-kernel times and the launch count come from the same step run eagerly (a graph replay makes the same launches
    # but bypasses the host-side event brackets)
    ops.Profiler.reset()
    ops.Profiler.enabled = True
    l0 = ops.Profiler.launches
    timed(lambda: trainer.eager_step(batch_d), args.steps)
    ops.Profiler.enabled = False
    launches = (ops.Profiler.launches - l0) // max(1, args.steps)
    prof = ops.Profiler.summary()
    trainer.finish()

    rays = n_rays * world
    value = rays * args.steps / (total_ms * 1e-3)
    e2e_value = rays * args.steps / (e2e_ms * 1e-3)
    mlp_ms = sum(prof.get(k, (0, 0.0))[1] for k in ("yn_mlp_fwd", "yn_mlp_bwd")) / args.steps
    flops = n_rays * (B.N_COARSE + B.N_COARSE + B.N_FINE) * B.FLOP_PER_POINT_TRAIN
    achieved = flops / (mlp_ms * 1e-3) / 1e12
    if rank == 0:
        line = dict(
            metric="train rays/sec (lego.yml step, 4096 rays/GPU, 64+192 points/ray)", value=round(value, 1), unit="rays/s",
            n_gpus=world, steps=args.steps, warmup=max(3, args.warmup), ms_per_step=round(total_ms / args.steps, 3),
            higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16 operands, f32 accumulate / master weights",
            data="synthetic",
            config={"workload": "lego.yml training step, 4096 rays/GPU, coarse+fine fwd/bwd + Adam, ray-sharded DDP",
                    "cuda_graph": use_graph,
                    "l2": "stash + gradient stash of one step (~10 GB) exceed the 126 MB L2", "parallelism": f"dp{world}"},
            clocks=clocks, gpu_launches=int(launches),
            e2e={"value": round(e2e_value, 1), "unit": "rays/s", "ms_per_step": round(e2e_ms / args.steps, 3),
                 "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            roofline={"bound": "tensor", "kernel": "mlp fwd + dgrad + wgrad kernels (coarse + fine)", "achieved": round(achieved, 1),
                      "peak": peaks["tensor"], "peak_source": f"{peaks['source']} bf16_tflops_sustained", "unit": "TFLOP/s",
                      "frac": round(achieved / peaks["tensor"], 4), "traffic": None,
                      "executed_note": "algorithmic FLOPs of SURVEY 8(d) (3 475 200 per point); the kernels issue 3 213 056: the "
                                       "linear intermediate layer is folded into the colour hidden layer (forward and data "
                                       "gradient), its weight gradient comes from a 128x256x256 post-product",
                      "kernel_ms_per_step": {k: round(v[1] / args.steps, 3) for k, v in prof.items()}},
        )
        B.emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
