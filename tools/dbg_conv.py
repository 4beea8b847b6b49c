import sys
sys.path.insert(0, "/root/repo/yet-another-nerf_b200"); sys.path.insert(0, "/root/repo")
import torch
from yanerf import synthetic as syn
from yanerf.testing import pipeline_cfg
from yanerf.pipelines import PIPELINES
from yanerf.pipelines.ray_samplers.utils import EvaluationMode
from yanerf.runners import FusedTrainer
DEV = torch.device("cuda")
for seed in range(6):
    torch.manual_seed(seed)
    cfg = pipeline_cfg(2, 2, 4, 5, 0.0, chunk=0, min_depth=0.1, max_depth=2.0)
    cfg.ray_sampler.n_pts_per_ray_training = 5
    cfg.ray_sampler.n_pts_per_ray_evaluation = 5
    pipe = PIPELINES.build(cfg).to(DEV)
    trainer = FusedTrainer(pipe, lr=5e-3)
    batch = dict(poses=syn.synth_camera(1, seed=0).to(DEV), focal_lengths=torch.full((1, 1), 2.0, device=DEV),
                 image_rgb=torch.rand(1, 2, 2, 3, device=DEV))
    hist = []
    for it in range(200):
        preds = trainer.train_step(batch)
        if it % 40 == 0: hist.append(round(float(preds["objective"].detach().mean()), 4))
    with torch.no_grad():
        ev = pipe(**batch, evaluation_mode=EvaluationMode.EVALUATION)
    print(seed, hist, "->", float(ev["objective"].mean()))
