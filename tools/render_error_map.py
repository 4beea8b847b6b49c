"""Diagnostic: kernel (fp16 operands) vs CPU oracle on reference-sized chunks spread over the 800x800 lego image
(corners, edges, centre) for two weight seeds: where does the rendered-RGB error peak?"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "yet-another-nerf_b200")); sys.path.insert(0, os.path.join(REPO, "tests"))
import torch
from oracle import nerf_oracle as O
from tools import synthetic as syn
from tools.testing import build_pipeline, load_synth_nets
from yanerf.pipelines.utils import EvaluationMode
from conftest import oracle_spec

torch.set_num_threads(os.cpu_count())
DEV = "cuda"
H = W = 800
for seeds in ((0, 1), (21, 22)):
    pipe = build_pipeline(H, W, 4096, 128, 0.2, 131072).to(DEV)
    nets = load_synth_nets(pipe, seeds=seeds, gain=1.0)
    poses, focal = syn.synth_camera(1, seed=0, jitter=0.0), torch.full((1, 1), syn.LEGO_FOCAL)
    with torch.no_grad():
        a = pipe(poses=poses.to(DEV), focal_lengths=focal.to(DEV), evaluation_mode=EvaluationMode.EVALUATION)
    img = a["rendered_images"].reshape(1, H * W, 3).cpu()
    per = 2045
    for chunk in (0, 1, 78, 150, 156, 234, 311, 312):
        s0 = chunk * per
        e0 = min(s0 + per, H * W)
        ref = O.render_image(nets, oracle_spec(H, W, 128, 0.2, 131072), poses, focal, ray_slice=(s0, e0))
        err = (img[:, s0:e0] - ref["features"]).abs()
        print(f"seeds {seeds} chunk {chunk:3d} (rows {s0 // W}-{e0 // W}): max abs {float(err.max()):.2e} mean {float(err.mean()):.2e}", flush=True)
