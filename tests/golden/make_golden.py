"""Generate the golden fixtures in this directory by EXECUTING THE UNMODIFIED
REFERENCE (/root/reference) on seeded inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The GPU box has no /root/reference; it only reads the committed *.npz files.
Inputs are regenerated from seeds by `tools/synthetic.py` (loaded by path so the
reference's own `yanerf` package stays the one on sys.path) or stored beside the
outputs when small.  Random draws are injected into the reference by replacing
torch.multinomial / rand_like / randn_like / rand with queues that replay
pre-generated tensors in the reference's call order.
"""
import contextlib
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

sys.path.insert(0, HERE)
import ref_shims  # noqa: E402

ref_shims.install()
sys.path.insert(0, REF)

spec = importlib.util.spec_from_file_location(
    "yn_synthetic", os.path.join(REPO, "tools", "synthetic.py")
)
syn = importlib.util.module_from_spec(spec)
spec.loader.exec_module(syn)

from yanerf.pipelines.builder import PIPELINES  # noqa: E402  (the reference's)
from yanerf.pipelines.models import MODELS  # noqa: E402
from yanerf.pipelines.ray_samplers.ray_sampler import _jiggle_within_stratas, _xy_to_ray_bundle  # noqa: E402
from yanerf.pipelines.ray_samplers.utils import get_xy_grid  # noqa: E402
from yanerf.pipelines.renderers.multipass_emission_absorpsion_renderer import EmissionAbsorptionRaymarcher  # noqa: E402
from yanerf.pipelines.renderers.utils import RayPointRefiner, sample_pdf_python  # noqa: E402
from yanerf.pipelines.utils import EvaluationMode  # noqa: E402
from yanerf.utils.config import Config, ConfigDict  # noqa: E402

assert sys.modules["yanerf"].__file__.startswith(REF), "must run against the reference package"


@contextlib.contextmanager
def inject(multinomial=(), rand_like=(), randn_like=(), rand=()):
    """Replay pre-generated draws in call order."""
    q = {"multinomial": list(multinomial), "rand_like": list(rand_like), "randn_like": list(randn_like), "rand": list(rand)}
    saved = {k: getattr(torch, k) for k in q}

    def mk(name):
        def f(*a, **kw):
            t = q[name].pop(0)
            return t.clone()
        return f

    for k in q:
        setattr(torch, k, mk(k))
    try:
        yield q
    finally:
        for k, v in saved.items():
            setattr(torch, k, v)
    for k, v in q.items():
        assert not v, f"unused injected draws for {k}"


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        out[k] = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}.npz: " + ", ".join(f"{k}{tuple(v.shape)}" for k, v in out.items()))


def rs(seed):
    return np.random.RandomState(seed)


def t32(a):
    return torch.from_numpy(np.asarray(a, dtype=np.float32))


# --------------------------------------------------------------------------- #
def golden_sample_pdf():
    r = rs(11)
    out = {}
    for tag, R, nb, ns in (("lego", 257, 63, 128), ("fern", 64, 63, 64), ("wide", 31, 191, 128)):
        z = np.sort(2 + 4 * r.uniform(size=(R, nb)).astype(np.float32), axis=-1)
        w = (r.uniform(size=(R, nb - 1)).astype(np.float32)) ** 4
        w[r.uniform(size=w.shape) < 0.3] = 0.0  # empty bins are the common case
        w[0] = 0.0  # an all-empty ray
        u = np.minimum(r.uniform(size=(R, ns)).astype(np.float32), np.float32(1 - 2**-24))
        bins, weights = t32(z), t32(w)
        det = sample_pdf_python(bins, weights, ns, det=True)
        with inject(rand=[t32(u)]):
            rnd = sample_pdf_python(bins, weights, ns, det=False)
        # indices: the reference's own lines 122-137 executed verbatim
        ww = weights + 1e-5
        pdf = ww / ww.sum(dim=-1, keepdim=True)
        cdf = torch.cumsum(pdf, -1)
        cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
        u_det = torch.linspace(0.0, 1.0, ns).expand(R, ns).contiguous()
        out.update({
            f"{tag}_bins": bins, f"{tag}_weights": weights, f"{tag}_u": t32(u),
            f"{tag}_det": det, f"{tag}_rnd": rnd, f"{tag}_cdf": cdf,
            f"{tag}_inds_det": torch.searchsorted(cdf, u_det, right=True),
            f"{tag}_inds_rnd": torch.searchsorted(cdf, t32(u), right=True),
        })
    save("sample_pdf", **out)


def golden_raymarch():
    r = rs(12)
    out = {}
    variants = {
        "lego": dict(background_density_bias=1e-6, blend_output=False, hard_background=False, bg_color=(0.0, 0.0, 0.0)),
        "blend": dict(background_density_bias=0.0, blend_output=True, hard_background=False, bg_color=(0.0, 0.0, 0.0)),
        "hard": dict(background_density_bias=0.0, blend_output=False, hard_background=True, bg_color=(0.3, 0.5, 0.7)),
    }
    for P in (64, 192):
        R = 33
        sig = t32(3 * r.standard_normal(size=(R, P, 1)) + 0.5)
        sig[1] = 0.0  # empty ray
        rgb = t32(r.uniform(size=(R, P, 3)))
        z = t32(np.sort(2 + 4 * r.uniform(size=(R, P)), axis=-1))
        d = t32(r.standard_normal(size=(R, 3)))
        noise = t32(r.standard_normal(size=(R, P)))
        bg = t32(r.uniform(size=(R, 3)))
        out.update({f"P{P}_sigma": sig, f"P{P}_rgb": rgb, f"P{P}_z": z, f"P{P}_d": d, f"P{P}_noise": noise, f"P{P}_bg": bg})
        for name, kw in variants.items():
            march = EmissionAbsorptionRaymarcher(surface_thickness=1, capping_function="exponential",
                                                 weight_function="product", background_opacity=1e10, **kw)
            for with_noise in (False, True):
                for with_bg in (False, True):
                    std = 0.2 if with_noise else 0.0
                    ctx = inject(randn_like=[noise]) if with_noise else contextlib.nullcontext()
                    with ctx:
                        f, dep, op, w, _ = march(sig, rgb, {}, z, d, density_noise_std=std, bg_color=bg if with_bg else None)
                    key = f"P{P}_{name}_n{int(with_noise)}_b{int(with_bg)}"
                    out.update({key + "_feat": f, key + "_depth": dep, key + "_opac": op, key + "_w": w})
    save("raymarch", **out)


MLP_CFGS = {
    "lego": dict(type="NeRFMLP", n_layers=8, input_skips=[5], n_harmonic_functions_xyz=10,
                 harmonic_functions_xyz_append_intput=True, n_hidden_neurons_xyz=256, n_harmonic_functions_dir=4,
                 harmonic_functions_dir_append_intput=True, n_hidden_neurons_dir=128, latent_dim=0, input_xyz=True,
                 input_dir=True, color_dim=3, nerf_paper_v1=False),
    "small": dict(type="NeRFMLP", n_layers=5, input_skips=[2], n_harmonic_functions_xyz=8,
                  harmonic_functions_xyz_append_intput=True, n_hidden_neurons_xyz=64, n_harmonic_functions_dir=4,
                  harmonic_functions_dir_append_intput=True, n_hidden_neurons_dir=32, latent_dim=0, input_xyz=True,
                  input_dir=True, color_dim=3),
    # the lego architecture with a 5-wide global code per image (nerf_mlp.py:158-170, 324-335); LAST, so that the random
    # stream of the earlier cases is unchanged
    "latent": dict(type="NeRFMLP", n_layers=8, input_skips=[5], n_harmonic_functions_xyz=10,
                   harmonic_functions_xyz_append_intput=True, n_hidden_neurons_xyz=256, n_harmonic_functions_dir=4,
                   harmonic_functions_dir_append_intput=True, n_hidden_neurons_dir=128, latent_dim=5, input_xyz=True,
                   input_dir=True, color_dim=3, nerf_paper_v1=False),
}


def load_synth(model, seed, gain):
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = syn.synth_mlp_state(shapes, seed, gain)
    model.load_state_dict(sd)
    return shapes


def golden_mlp():
    r = rs(13)
    out = {}
    for name, cfg in MLP_CFGS.items():
        for gain, seed in ((1.0, 7), (3.0, 8)):
            model = MODELS.build(dict(cfg))
            load_synth(model, seed, gain)
            B, n, P = 2, 9, 24
            o = t32(r.uniform(-0.2, 0.2, size=(B, n, 1, 3)) + np.array([0, 0, -4.0]))
            d = t32(r.uniform(-0.4, 0.4, size=(B, n, 1, 3)) + np.array([0, 0, 1.0]))
            z = t32(np.sort(2 + 4 * r.uniform(size=(B, n, 1, P)), axis=-1))
            key = f"{name}_g{int(gain)}"
            codes = None
            if cfg.get("latent_dim", 0) > 0:
                codes = t32(r.standard_normal(size=(B, cfg["latent_dim"])))
                out[key + "_codes"] = codes
            with torch.no_grad():
                res = model(o, d, z, global_codes=codes)
            out.update({key + "_o": o, key + "_d": d, key + "_z": z,
                        key + "_density": res["rays_densities"], key + "_rgb": res["rays_features"]})
    save("mlp", **out)


def golden_refiner():
    r = rs(14)
    out = {}
    for tag, P, N in (("lego", 64, 128), ("fern", 64, 64)):
        R = 40
        z = t32(np.sort(2 + 4 * r.uniform(size=(1, R, 1, P)), axis=-1))
        w = t32(r.uniform(size=(1, R, 1, P)) ** 6)
        u = t32(np.minimum(r.uniform(size=(R, N)).astype(np.float32), np.float32(1 - 2**-24)))
        o = torch.zeros(1, R, 1, 3)
        det = RayPointRefiner(N, random_sampling=False)(o, o, z, o[..., :2], w).lengths
        with inject(rand=[u]):
            rnd = RayPointRefiner(N, random_sampling=True)(o, o, z, o[..., :2], w).lengths
        no_add = RayPointRefiner(N, random_sampling=False, add_input_samples=False)(o, o, z, o[..., :2], w).lengths
        out.update({f"{tag}_z": z, f"{tag}_w": w, f"{tag}_u": u, f"{tag}_det": det, f"{tag}_rnd": rnd, f"{tag}_noadd": no_add})
    save("refiner", **out)


def golden_raysampler():
    r = rs(15)
    B, H, W, P = 2, 6, 10, 16
    poses = syn.synth_camera(B, seed=3, jitter=0.3)
    focal = t32([[11.5], [13.25]])
    grid = get_xy_grid(H, W)[None].expand(B, -1, -1, -1)
    ev = _xy_to_ray_bundle(poses, W, H, focal, grid, 0.5, 3.0, P, False)
    u = t32(np.minimum(r.uniform(size=(B, 7, 1, P)).astype(np.float32), np.float32(1 - 2**-24)))
    pix = torch.from_numpy(np.stack([r.choice(H * W, 7, replace=False) for _ in range(B)]).astype(np.int64))
    xy = torch.gather(grid.reshape(B, -1, 2), 1, pix[..., None].expand(-1, -1, 2))[:, :, None]
    with inject(rand_like=[u]):
        tr = _xy_to_ray_bundle(poses, W, H, focal, xy, 0.5, 3.0, P, True)
    zlin = torch.linspace(0.5, 3.0, P)[None].expand(5, -1)
    u2 = t32(r.uniform(size=(5, P)))
    with inject(rand_like=[u2]):
        jig = _jiggle_within_stratas(zlin)
    save("raysampler", poses=poses, focal=focal, pix=pix, u=u, u2=u2, jig=jig,
         ev_origins=ev.origins, ev_directions=ev.directions, ev_lengths=ev.lengths, ev_xys=ev.xys,
         tr_origins=tr.origins, tr_directions=tr.directions, tr_lengths=tr.lengths, tr_xys=tr.xys)


def pipeline_cfg(arch, H, W, n_rays, n_fine, noise_std, chunk):
    return ConfigDict(dict(
        type="NeRFPipeline", chunk_size_grid=chunk, num_passes=2, output_rasterized_mc=True,
        loss_weights={"loss_prev_stage_rgb_mse": 1.0, "loss_rgb_mse": 1.0},
        model=dict(MLP_CFGS[arch]),
        ray_sampler=dict(type="RaySampler", image_height=H, image_width=W, min_depth=2.0, max_depth=6.0,
                         n_pts_per_ray_evaluation=64, n_pts_per_ray_training=64,
                         n_rays_per_image_sampled_from_mask=n_rays, scene_extent=0.0,
                         stratified_point_sampling_training=True, stratified_point_sampling_evaluation=False),
        renderer=dict(type="MultipassEmissionAbsorpsionRenderer", append_coarse_samples_to_fine=True,
                      bg_color=[0.0, 0.0, 0.0], blend_output=False, density_noise_std_train=noise_std,
                      n_pts_per_ray_fine_evaluation=n_fine, n_pts_per_ray_fine_training=n_fine,
                      hard_background=False, background_density_bias=1.0e-6),
        feature_extractor=[],
    ))


def golden_pipeline():
    out = {}
    # "lego": stress weights (x3: raw densities O(1e3), saturated sigmoids); "lego1": the same lego.yml configuration
    # (64+128 samples, density noise 0.2) at xavier scale, where a max-abs comparison with a 16-bit implementation is
    # meaningful; "fern": 64+64 samples, no noise
    for tag, n_fine, std, gain in (("lego", 128, 0.2, 3.0), ("fern", 64, 0.0, 1.0), ("lego1", 128, 0.2, 1.0)):
        B, H, W, n = 2, 16, 20, 48
        cfg = pipeline_cfg("lego", H, W, n, n_fine, std, chunk=64 * 37)
        pipe = PIPELINES.build(cfg)
        for k, fn in enumerate(pipe.implicit_functions):
            load_synth(fn._fn, 21 + k, gain)
        poses = syn.synth_camera(B, seed=5)
        focal = torch.full((B, 1), 25.0)
        image = syn.synth_image(B, H, W, seed=4)
        dr = syn.synth_draws(B, n, H * W, 64, n_fine, seed=6)
        randn = [dr["noise0"].view(B, n, 1, 64), dr["noise1"].view(B, n, 1, 64 + n_fine)] if std > 0 else []
        with inject(multinomial=[dr["pix"]], rand_like=[dr["u_strat"].view(B, n, 1, 64)], randn_like=randn, rand=[dr["u_pdf"]]):
            preds = pipe(poses=poses, focal_lengths=focal, image_rgb=image, evaluation_mode=EvaluationMode.TRAINING)
        preds["objective"].mean().backward()
        for k in ("loss_rgb_mse", "loss_rgb_huber", "loss_prev_stage_rgb_mse", "loss_prev_stage_rgb_huber", "objective",
                  "rendered_images", "rendered_depths", "rendered_alpha_masks"):
            out[f"{tag}_train_{k}"] = preds[k]
        for k, fn in enumerate(pipe.implicit_functions):
            for name, p in fn._fn.named_parameters():
                g = p.grad.reshape(-1)
                out[f"{tag}_grad{k}_{name}"] = torch.stack([g.sum(), g.abs().sum(), g.norm(), *g[:5], *g[-5:]])
            # one full small gradient for element-wise comparison
            out[f"{tag}_grad{k}_full_density_w"] = fn._fn.density_layer.weight.grad
            out[f"{tag}_grad{k}_full_color2_w"] = fn._fn.color_layer[2].weight.grad
            out[f"{tag}_grad{k}_full_l0_b"] = fn._fn.xyz_encoder.mlp[0][0].bias.grad
        with torch.no_grad():
            ev = pipe(poses=poses, focal_lengths=focal, image_rgb=image, evaluation_mode=EvaluationMode.EVALUATION)
        for k in ("loss_rgb_mse", "loss_prev_stage_rgb_mse", "objective", "rendered_images", "rendered_depths", "rendered_alpha_masks"):
            out[f"{tag}_eval_{k}"] = ev[k]
    save("pipeline", **out)


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(8)
    golden_sample_pdf()
    golden_raymarch()
    golden_mlp()
    golden_refiner()
    golden_raysampler()
    golden_pipeline()
