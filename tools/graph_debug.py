"""Capture the training step of the small test pipeline under a CUDA graph, in isolation (debug aid)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "yet-another-nerf_b200"))
import torch
from yanerf import synthetic as syn
from yanerf.testing import pipeline_cfg
from yanerf.pipelines import PIPELINES
from yanerf.runners import FusedTrainer

DEV = torch.device("cuda:0")
H = W = 24
img = torch.rand(1, H, W, 3)
for npts in (32, 64):
    cfg = pipeline_cfg(H, W, 256, npts, 0.0, chunk=131072)
    cfg.ray_sampler.n_pts_per_ray_training = npts
    pipe = PIPELINES.build(cfg).to(DEV)
    tr = FusedTrainer(pipe, lr=5e-4, use_cuda_graph=True)
    batch = dict(poses=syn.synth_camera(1, seed=0, jitter=0.0).to(DEV), focal_lengths=torch.full((1, 1), 30.0, device=DEV), image_rgb=img.to(DEV))
    try:
        for it in range(8):
            p = tr.train_step(batch)
        torch.cuda.synchronize()
        print(npts, "ok", float(p["objective"].mean()))
    except Exception as e:
        print(npts, "FAILED", str(e)[:200])
