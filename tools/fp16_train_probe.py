"""Probe: does the lego-architecture pipeline train with fp16 tensor-core operands for activations AND gradients
(no loss scaling), compared with the bf16 default?  Prints PSNR after 400 fused steps on the 24x24 two-tone image and
how many colour-head / density gradient entries fall below fp16's normal / subnormal range."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "yet-another-nerf_b200"))
import torch
from tools import synthetic as syn
from tools.testing import pipeline_cfg
from yanerf.pipelines import PIPELINES
from yanerf.pipelines.utils import EvaluationMode
from yanerf.runners import FusedTrainer
from yanerf.runners.apis import create_stats

DEV = "cuda"
H = W = 24
yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
img = torch.where(((xx - 12) ** 2 + (yy - 12) ** 2 < 49)[..., None], torch.tensor([0.9, 0.2, 0.1]), torch.tensor([0.1, 0.3, 0.8]))
for dtype in ("bf16", "fp16"):
    torch.manual_seed(1)
    cfg = pipeline_cfg(H, W, 256, 32, 0.0, chunk=131072)
    cfg.ray_sampler.n_pts_per_ray_training = 32
    cfg.ray_sampler.n_pts_per_ray_evaluation = 32
    pipe = PIPELINES.build(cfg).to(DEV)
    for fn in pipe.implicit_functions:
        fn._fn.set_operand_dtype(dtype, training=True)
    trainer = FusedTrainer(pipe, lr=5e-4)
    batch = dict(poses=syn.synth_camera(1, seed=0, jitter=0.0).to(DEV), focal_lengths=torch.full((1, 1), 30.0, device=DEV),
                 image_rgb=img[None].to(DEV))
    for it in range(400):
        preds = trainer.train_step(batch)
        if it in (0, 399):
            g = trainer.flat_grad
            print(dtype, "step", it, "objective", float(preds["objective"].mean()), "|grad| max", float(g.abs().max()),
                  "frac zero", float((g == 0).float().mean()))
    trainer.finish()
    with torch.no_grad():
        ev = pipe(**batch, evaluation_mode=EvaluationMode.EVALUATION)
    print(dtype, "PSNR after 400 steps:", create_stats(ev)["loss_rgb_psnr"])
