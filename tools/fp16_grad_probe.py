"""Probe: gradients of one lego.yml training step (4096 rays, mean-reduced losses: d(loss)/d(activation) ~ 1e-6 .. 1e-9) with
bf16 operands vs fp16 operands + the per-call power-of-two gradient scale, against each other: norms, zero fractions, cosine."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "yet-another-nerf_b200"))
import torch
import bench as B
from yanerf.pipelines.utils import EvaluationMode
from yanerf import ops

dev = torch.device("cuda")
grads = {}
for dtype in ("bf16", "fp16"):
    pipe, _ = B.build_lego_pipeline(dev)
    for fn in pipe.implicit_functions:
        fn._fn.set_operand_dtype(dtype, training=True)
    rng = ops.DeviceRng(dev, seed=5)
    pipe.set_device_rng(rng)
    pipe.ray_sampler.fused_pixel_sampler = True
    ops.step_begin(rng, None)
    poses, focal, image = (t.to(dev) for t in B.synthetic_inputs(0))
    out = pipe(poses=poses, focal_lengths=focal, image_rgb=image, evaluation_mode=EvaluationMode.TRAINING)
    out["objective"].mean().backward()
    g = torch.cat([p.grad.reshape(-1) for p in pipe.parameters()])
    grads[dtype] = g
    print(dtype, "objective", float(out["objective"].mean()), "|g|", float(g.norm()), "max", float(g.abs().max()),
          "zero frac", float((g == 0).float().mean()), "finite", bool(torch.isfinite(g).all()))
a, b = grads["bf16"].double(), grads["fp16"].double()
print("cos(bf16, fp16)", float((a * b).sum() / (a.norm() * b.norm())), "norm ratio", float(b.norm() / a.norm()))
