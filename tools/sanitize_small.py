"""Smallest end-to-end case for compute-sanitizer: two fused training steps (in-kernel draws, fused rays / loss / scatter
kernels, fp16 gradient scaling, Adam) + one tiny eval render + the torch-generator path of the same pipeline."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "yet-another-nerf_b200"))
import torch
from tools import synthetic as syn
from yanerf.pipelines.utils import EvaluationMode
from yanerf.runners import FusedTrainer
from tools.testing import build_pipeline, load_synth_nets
dev = "cuda"
pipe = build_pipeline(8, 8, 24, 64, 0.2, 64 * 9).to(dev)
load_synth_nets(pipe, (1, 2), 1.0)
poses, focal, img = syn.synth_camera(1, 0).to(dev), torch.full((1, 1), 10.0, device=dev), syn.synth_image(1, 8, 8, 3).to(dev)
out = pipe(poses=poses, focal_lengths=focal, image_rgb=img, evaluation_mode=EvaluationMode.TRAINING)  # torch draws
out["objective"].mean().backward()
trainer = FusedTrainer(pipe, lr=1e-3)  # in-kernel draws, flat buffers
for _ in range(2):
    preds = trainer.train_step(dict(poses=poses, focal_lengths=focal, image_rgb=img))
trainer.finish()
with torch.no_grad():
    ev = pipe(poses=poses, focal_lengths=focal, image_rgb=img, evaluation_mode=EvaluationMode.EVALUATION)
torch.cuda.synchronize()
print("ok", float(out["objective"].mean()), float(preds["objective"].mean()), float(ev["objective"].mean()))
