"""Golden fixtures for the pipeline's OPTIONAL inputs and renderer switches, from the UNMODIFIED reference:

    python tests/golden/make_golden_variants.py      (build container only; writes pipeline_variants.npz)

Cases (all 2 images of 16x20, 48 rays, 64+64 samples, xavier-scale synthetic weights; draws replayed like make_golden.py):
  bgdepth  blend_output=True, per-pixel background image (`bg_image_rgb`) and a depth map (`depth_map` -> loss_depth_abs)
  hardbg   hard_background=True with a white constant background
  mask     MASK_SAMPLE training with a `sampling_prob_mask` (the pick itself is replayed)
  custom   evaluation at a custom `image_height` x `image_width`, `min_depth` / `max_depth` overrides
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (installs the shims, puts the reference on sys.path)

from yanerf.pipelines.builder import PIPELINES  # noqa: E402
from yanerf.pipelines.utils import EvaluationMode  # noqa: E402

syn = MG.syn
EVAL_KEYS = ("rendered_images", "rendered_depths", "rendered_alpha_masks")


def build(renderer_over=None, n_fine=64, std=0.0, H=16, W=20, n=48):
    cfg = MG.pipeline_cfg("lego", H, W, n, n_fine, std, chunk=64 * 37)
    cfg["renderer"].update(renderer_over or {})
    pipe = PIPELINES.build(cfg)
    for k, fn in enumerate(pipe.implicit_functions):
        MG.load_synth(fn._fn, 31 + k, 1.0)
    return pipe


def grads_summary(pipe, out, tag):
    for k, fn in enumerate(pipe.implicit_functions):
        for name, p in fn._fn.named_parameters():
            g = p.grad.reshape(-1)
            out[f"{tag}_grad{k}_{name}"] = torch.stack([g.sum(), g.abs().sum(), g.norm()])
            p.grad = None


def keep(out, tag, preds):
    for k, v in preds.items():
        if torch.is_tensor(v):
            out[f"{tag}_{k}"] = v.detach()


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    out = {}
    B, H, W, n, n_fine = 2, 16, 20, 48, 64
    poses, focal = syn.synth_camera(B, seed=8), torch.full((B, 1), 25.0)
    image = syn.synth_image(B, H, W, seed=9)
    bg_image = syn.synth_image(B, H, W, seed=10)
    depth_map = 2.0 + 4.0 * syn.synth_image(B, H, W, seed=11)[..., :1]
    dr = syn.synth_draws(B, n, H * W, 64, n_fine, seed=12)
    draws = lambda: dict(multinomial=[dr["pix"]], rand_like=[dr["u_strat"].view(B, n, 1, 64)], rand=[dr["u_pdf"]])

    # ---- bgdepth
    pipe = build(dict(blend_output=True))
    with MG.inject(**draws()):
        preds = pipe(poses=poses, focal_lengths=focal, image_rgb=image, bg_image_rgb=bg_image, depth_map=depth_map,
                     evaluation_mode=EvaluationMode.TRAINING)
    preds["objective"].mean().backward()
    keep(out, "bgdepth_train", preds)
    grads_summary(pipe, out, "bgdepth")
    with torch.no_grad():
        ev = pipe(poses=poses, focal_lengths=focal, image_rgb=image, bg_image_rgb=bg_image, depth_map=depth_map,
                  evaluation_mode=EvaluationMode.EVALUATION)
    keep(out, "bgdepth_eval", ev)

    # ---- hardbg
    pipe = build(dict(hard_background=True, bg_color=[1.0, 1.0, 1.0]))
    with MG.inject(**draws()):
        preds = pipe(poses=poses, focal_lengths=focal, image_rgb=image, evaluation_mode=EvaluationMode.TRAINING)
    preds["objective"].mean().backward()
    keep(out, "hardbg_train", preds)
    grads_summary(pipe, out, "hardbg")
    with torch.no_grad():
        ev = pipe(poses=poses, focal_lengths=focal, image_rgb=image, evaluation_mode=EvaluationMode.EVALUATION)
    keep(out, "hardbg_eval", ev)

    # ---- mask
    pipe = build()
    # (`mask_crop` cannot be exercised: the reference's mask branch, ray_sampler.py:82-97, reads a non-existent
    # `self.image_height` when no size is passed and an unbound `_image_height` when one is)
    prob = syn.synth_image(B, H, W, seed=14)[..., 1] + 0.1
    out["mask_prob"] = prob
    with MG.inject(**draws()):
        preds = pipe(poses=poses, focal_lengths=focal, image_rgb=image, sampling_prob_mask=prob,
                     evaluation_mode=EvaluationMode.TRAINING)
    keep(out, "mask_train", preds)

    # ---- custom grid and depth range at evaluation time
    with torch.no_grad():
        ev = pipe(poses=poses, focal_lengths=focal, image_rgb=image, image_height=12, image_width=10, min_depth=1.5,
                  max_depth=5.0, evaluation_mode=EvaluationMode.EVALUATION)
    keep(out, "custom_eval", ev)
    MG.save("pipeline_variants", **out)


if __name__ == "__main__":
    main()
