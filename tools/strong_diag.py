"""Where the strong-scaling render loses time: per-rank split of one sharded 800x800 render into
[start -> before the all-gather] and [all-gather -> end], plus the fixed cost of a 16x16 sharded render.
(Rate-proportional slabs were tried with this tool and lost: the per-GPU rates jitter by 2-3 % from render to render,
so a split measured on earlier renders is as often wrong as right)

torchrun --nproc-per-node N tools/strong_diag.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from yanerf.pipelines import nerf_pipeline as npl
    from yanerf.pipelines.utils import EvaluationMode
    from yanerf.runners.apis import enable_ray_sharding

    pipe, _ = bench.build_lego_pipeline(dev)
    poses, focal, image = bench.synthetic_inputs(0)
    batch = dict(poses=poses.to(dev), focal_lengths=focal.to(dev), image_rgb=image.to(dev))

    marks = {}
    orig = npl.gather_slabs

    def patched(*a, **k):
        marks["pre"] = torch.cuda.Event(enable_timing=True)
        marks["pre"].record()
        return orig(*a, **k)

    npl.gather_slabs = patched

    def step(**extra):
        with torch.no_grad():
            return pipe(**batch, evaluation_mode=EvaluationMode.EVALUATION, **extra)

    med = lambda v: sorted(v)[len(v) // 2]

    def segment(name, n, extra):
        loc, tail, tot = [], [], []
        for _ in range(n):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step(**extra)
            e1.record()
            torch.cuda.synchronize()
            loc.append(e0.elapsed_time(marks["pre"]))
            tail.append(marks["pre"].elapsed_time(e1))
            tot.append(e0.elapsed_time(e1))
        print(f"[{name}] rank {rank}: local {med(loc):.3f} ms, gather+tail {med(tail):.3f} ms, total {med(tot):.3f} ms "
              f"(min {min(tot):.3f}, max {max(tot):.3f})", flush=True)

    enable_ray_sharding(pipe)
    for _ in range(6):
        step()
    for rnd in range(3):
        segment(f"800x800 r{rnd}", 12, {})
    segment("16x16", 12, dict(image_height=16, image_width=16))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
