"""Timeline of ONE ray-slab sharded 800x800 render (kineto: host ops + every kernel) -> gpurun_out/strong_trace_rank0.json.

torchrun --nproc-per-node N tools/strong_trace.py
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
    from yanerf.pipelines.utils import EvaluationMode
    from yanerf.runners.apis import enable_ray_sharding

    pipe, _ = bench.build_lego_pipeline(dev)
    enable_ray_sharding(pipe)
    poses, focal, image = bench.synthetic_inputs(0)
    batch = dict(poses=poses.to(dev), focal_lengths=focal.to(dev), image_rgb=image.to(dev))

    def step():
        with torch.no_grad():
            return pipe(**batch, evaluation_mode=EvaluationMode.EVALUATION)

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            step()
        torch.cuda.synchronize()
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        prof.export_chrome_trace("gpurun_out/strong_trace_rank0.json")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
