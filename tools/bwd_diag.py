import sys, os
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "yet-another-nerf_b200")); sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np, torch
from oracle import nerf_oracle as O
from tools import synthetic as syn
from yanerf.pipelines.models import MODELS
from tools.testing import LEGO_MLP
DEV = "cuda"
T = lambda a: torch.from_numpy(np.asarray(a))
for dtype in ("bf16", "fp16"):
    spec = O.MLPSpec()
    mlp = MODELS.build(dict(LEGO_MLP)); mlp.set_operand_dtype(dtype)
    sd = syn.synth_mlp_state(spec.param_shapes(), 13, 1.0); mlp.load_state_dict(sd); mlp = mlp.to(DEV)
    R, P = 70, 64
    rs = np.random.RandomState(R * 7 + P)
    o = T((rs.uniform(-0.2, 0.2, size=(R, 3)) + np.array([0, 0, -4.0])).astype(np.float32))
    d = T((rs.uniform(-0.4, 0.4, size=(R, 3)) + np.array([0, 0, 1.0])).astype(np.float32))
    z = T(np.sort(2 + 4 * rs.uniform(size=(R, P)), axis=-1).astype(np.float32))
    gd = T(rs.standard_normal(size=(R, P)).astype(np.float32)); gc = T(rs.standard_normal(size=(R, P, 3)).astype(np.float32))
    for which in ("both", "density_only", "rgb_only"):
        ref_params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        dens_ref, rgb_ref = O.mlp_forward(ref_params, spec, o, d, z)
        a = 0.0 if which == "rgb_only" else 1.0; b = 0.0 if which == "density_only" else 1.0
        (a * (dens_ref * gd).sum() + b * (rgb_ref * gc).sum()).backward()
        mlp.zero_grad()
        out = mlp(o.to(DEV)[None], d.to(DEV)[None], z.to(DEV)[None])
        (a * (out["rays_densities"][0, ..., 0] * gd.to(DEV)).sum() + b * (out["rays_features"][0] * gc.to(DEV)).sum()).backward()
        print(f"== {dtype} {which}")
        for k, p in mlp.named_parameters():
            g, gr = p.grad.detach().cpu().double().reshape(-1), ref_params[k].grad.double().reshape(-1)
            rel = float((g - gr).norm() / gr.norm().clamp_min(1e-12)); cos = float(torch.dot(g, gr) / (g.norm() * gr.norm()).clamp_min(1e-30))
            print(f"  {k:38s} rel {rel:.3e} cos {cos:.6f} |ref| {float(gr.norm()):.3e} |got| {float(g.norm()):.3e}")
