"""Records (a) the on-disk layout of a REFERENCE checkpoint (scripts/run.py:416-422: {"model": state_dict, "optimizer":
torch.optim.Adam.state_dict(), "epoch": e}) as metadata only: ordered state-dict keys + shapes, the parameter order the
optimizer indexes by, and the optimizer state-dict structure with tensors replaced by their shapes -- the optimizer is built the way scripts/run.py
builds it: `create_param_groups` (runners/utils.py:142-186, every group carries `init_lr`) + `torch.optim.Adam`
(run.py:158-159); and (b) the per-iteration learning rates the reference's schedulers set (`create_lr_scheduler` +
`warmup_lr_scheduler`, runners/utils.py:65-109, called as in runners/apis.py:77-79) for the lego.yml / fern.yml runner
values at world sizes 1 and 8 and for a cosine variant -> lr_schedule.json.  Run in the build container only (needs
/root/reference):
    python tests/golden/make_checkpoint_fixture.py
The GPU box / the tests read the committed checkpoint_layout.json."""
import ast
import json
import logging
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402

ref_shims.install()
ast.Str = getattr(ast, "Str", str)  # runners/utils.py:3 imports ast.Str (a type annotation; gone in Python 3.12)
sys.path.insert(0, REF)
from yanerf.pipelines.builder import PIPELINES  # noqa: E402  (the reference's)
from yanerf.pipelines.utils import EvaluationMode  # noqa: E402
from yanerf.runners.utils import create_lr_scheduler, create_param_groups, warmup_lr_scheduler  # noqa: E402
from yanerf.utils.config import ConfigDict  # noqa: E402

assert sys.modules["yanerf"].__file__.startswith(REF), "must run against the reference package"

# configs/nerf/lego.yml:45-94 at a tiny image size
MLP = dict(type="NeRFMLP", n_layers=8, input_skips=[5], n_harmonic_functions_xyz=10,
           harmonic_functions_xyz_append_intput=True, n_hidden_neurons_xyz=256, n_harmonic_functions_dir=4,
           harmonic_functions_dir_append_intput=True, n_hidden_neurons_dir=128, latent_dim=0, input_xyz=True,
           input_dir=True, color_dim=3, nerf_paper_v1=False)
cfg = ConfigDict(dict(
    type="NeRFPipeline", chunk_size_grid=4096, num_passes=2, output_rasterized_mc=True,
    loss_weights={"loss_prev_stage_rgb_mse": 1.0, "loss_rgb_mse": 1.0}, model=MLP,
    ray_sampler=dict(type="RaySampler", image_height=8, image_width=8, min_depth=2.0, max_depth=6.0,
                     n_pts_per_ray_evaluation=8, n_pts_per_ray_training=8, n_rays_per_image_sampled_from_mask=16,
                     scene_extent=0.0, stratified_point_sampling_training=True, stratified_point_sampling_evaluation=False),
    renderer=dict(type="MultipassEmissionAbsorpsionRenderer", append_coarse_samples_to_fine=True, bg_color=[0.0, 0.0, 0.0],
                  blend_output=False, density_noise_std_train=0.0, n_pts_per_ray_fine_evaluation=8,
                  n_pts_per_ray_fine_training=8, hard_background=False, background_density_bias=1.0e-6),
    feature_extractor=[],
))
torch.manual_seed(0)
model = PIPELINES.build(cfg)
# runner section of configs/nerf/lego.yml:12-43 (fern.yml has the same values)
RUNNER = dict(init_lr=5.0e-4, weight_decay=0.0, warmup_steps=1000, warmup_lr=1.0e-5, linear_scale=True,
              lr_decay_type="exponential", min_lr=5.0e-5, lr_decay_rate=0.1, lr_decay_iters=250000, num_iters=200000,
              lr_param_groups=[])
log = logging.getLogger("fixture")
runner = ConfigDict(dict(RUNNER))
optimizer = torch.optim.Adam(create_param_groups(model, runner, log), lr=runner.init_lr,
                             weight_decay=runner.weight_decay)  # scripts/run.py:158-159
pose = torch.eye(4)[None, :3].clone()
pose[0, 2, 3] = -4.0
for _ in range(2):
    out = model(poses=pose, focal_lengths=torch.full((1, 1), 10.0), image_rgb=torch.rand(1, 8, 8, 3),
                evaluation_mode=EvaluationMode.TRAINING)
    optimizer.zero_grad()
    out["objective"].mean().backward()
    optimizer.step()
ckpt = {"model": model.state_dict(), "optimizer": optimizer.state_dict(), "epoch": 3}  # run.py:416-422


def describe(o):
    if torch.is_tensor(o):
        return {"tensor": list(o.shape), "dtype": str(o.dtype).replace("torch.", "")}
    if isinstance(o, dict):
        return {str(k): describe(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [describe(v) for v in o]
    return o


layout = {
    "generated_by": "tests/golden/make_checkpoint_fixture.py against /root/reference, torch " + torch.__version__,
    "model_keys": [[k, list(v.shape)] for k, v in ckpt["model"].items()],
    "parameter_order": [n for n, _ in model.named_parameters()],
    "optimizer": describe(ckpt["optimizer"]),
    "epoch": ckpt["epoch"],
}
json.dump(layout, open(os.path.join(HERE, "checkpoint_layout.json"), "w"), indent=1)


def lr_trajectory(runner_values, world, iters):
    """The learning rate the reference's training loop sets at iteration `it` (runners/apis.py:77-79), with the
    world-size scaling of scripts/run.py:152-156 applied to init_lr / min_lr first."""
    cfg = ConfigDict(dict(runner_values))
    if world > 1 and cfg.linear_scale:
        cfg.init_lr = cfg.init_lr * world
        cfg.min_lr = cfg.min_lr * world
    net = torch.nn.Linear(2, 2)
    opt = torch.optim.Adam(create_param_groups(net, cfg, log), lr=cfg.init_lr, weight_decay=cfg.weight_decay)
    sched = create_lr_scheduler(opt, cfg)
    out = []
    for it in iters:
        sched(iter=it)
        if cfg["warmup_steps"] > 0 and it <= cfg["warmup_steps"]:
            warmup_lr_scheduler(opt, it, cfg["warmup_steps"], cfg["warmup_lr"])
        out.append(opt.param_groups[0]["lr"])
    return out


ITERS = [0, 1, 2, 10, 499, 500, 999, 1000, 1001, 1002, 5000, 50000, 100000, 199999, 200000, 250000, 300000, 400000]
cases = {
    "lego_world1": (RUNNER, 1),
    "lego_world8": (RUNNER, 8),
    "cosine_world2": ({**RUNNER, "lr_decay_type": "cosine", "lr_decay_iters": 1, "warmup_steps": 0}, 2),
    "no_scale_world4": ({**RUNNER, "linear_scale": False, "warmup_steps": 100, "warmup_lr": 0.0}, 4),
}
lr_fixture = {"generated_by": layout["generated_by"], "iters": ITERS,
              "cases": {k: {"runner": r, "world": w, "lr": lr_trajectory(r, w, ITERS)} for k, (r, w) in cases.items()}}
json.dump(lr_fixture, open(os.path.join(HERE, "lr_schedule.json"), "w"), indent=1)
print("lego lr @200k:", lr_fixture["cases"]["lego_world1"]["lr"][ITERS.index(200000)])
print(len(layout["model_keys"]), "state-dict entries,", len(layout["parameter_order"]), "parameters")
print(json.dumps(layout["optimizer"]["param_groups"])[:600])
print(json.dumps(layout["optimizer"]["state"]["0"]))
