#!/usr/bin/env python
"""torchrun check of the ray-slab sharded render: every rank renders its slab of ONE 800x800 image,
one all-gather per stage rebuilds it; compared bit for bit with rank 0's unsharded render, and timed.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/render_sharded_check.py"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "yet-another-nerf_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench as B  # noqa: E402
from yanerf.pipelines.utils import EvaluationMode  # noqa: E402
from yanerf.runners.apis import enable_ray_sharding  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
pipe, _ = B.build_lego_pipeline(dev)
poses, focal, image = B.synthetic_inputs(0)  # the same image on every rank
batch = dict(poses=poses.to(dev), focal_lengths=focal.to(dev), image_rgb=image.to(dev))


def timed(n=3):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        out = pipe(**batch, evaluation_mode=EvaluationMode.EVALUATION)
    e.record()
    torch.cuda.synchronize()
    return out, s.elapsed_time(e) / n


with torch.no_grad():
    pipe(**batch, evaluation_mode=EvaluationMode.EVALUATION)
    ref, ms_full = timed()
    sharded = enable_ray_sharding(pipe)
    pipe(**batch, evaluation_mode=EvaluationMode.EVALUATION)
    out, ms_shard = timed()
same = all(torch.equal(out[k], ref[k]) for k in ("rendered_images", "rendered_depths", "rendered_alpha_masks"))
t = torch.tensor([ms_shard], device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"check": "ray-slab sharded render", "world": world, "sharded": sharded, "bit_identical": same,
                      "ms_unsharded": round(ms_full, 2), "ms_sharded_max_over_ranks": round(float(t), 2),
                      "rays_per_s_sharded": round(640000 / (float(t) * 1e-3)),
                      "psnr": float(-10 * torch.log10(out["loss_rgb_mse"].mean()))}))
assert same
if world > 1:
    dist.destroy_process_group()
