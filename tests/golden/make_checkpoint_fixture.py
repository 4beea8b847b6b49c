"""Records the on-disk layout of a REFERENCE checkpoint (scripts/run.py:416-422: {"model": state_dict, "optimizer":
torch.optim.Adam.state_dict(), "epoch": e}) as metadata only: ordered state-dict keys + shapes, the parameter order the
optimizer indexes by, and the optimizer state-dict structure with tensors replaced by their shapes.  Run in the build
container only (needs /root/reference):
    python tests/golden/make_checkpoint_fixture.py
The GPU box / the tests read the committed checkpoint_layout.json."""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402

ref_shims.install()
sys.path.insert(0, REF)
from yanerf.pipelines.builder import PIPELINES  # noqa: E402  (the reference's)
from yanerf.pipelines.utils import EvaluationMode  # noqa: E402
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
optimizer = torch.optim.Adam(model.parameters(), lr=5e-4)  # scripts/run.py:159
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
print(len(layout["model_keys"]), "state-dict entries,", len(layout["parameter_order"]), "parameters")
print(json.dumps(layout["optimizer"]["param_groups"])[:600])
print(json.dumps(layout["optimizer"]["state"]["0"]))
