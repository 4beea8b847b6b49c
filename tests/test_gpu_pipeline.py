"""Pipeline-level GPU parity: the lego.yml-shaped NeRFPipeline against the golden outputs of the unmodified
reference (tests/golden/pipeline.npz) and the reference's own known-answer tests
(/root/reference/tests/test_pipeline.py:16-29,128-151, test_ray_sampler.py:27-99, test_models.py:19-55)."""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O
from tools import synthetic as syn
from yanerf.pipelines.utils import EvaluationMode, sample_grid, scatter_rays_to_image
from conftest import oracle_spec
from tools.testing import build_pipeline, load_synth_nets, pipeline_cfg
from yanerf.utils.config import ConfigDict

pytestmark = pytest.mark.gpu
DEV = "cuda"


def T(a):
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize("tag,n_fine,std,gain", [("fern", 64, 0.0, 1.0), ("lego1", 128, 0.2, 1.0), ("lego", 128, 0.2, 3.0)])
@pytest.mark.parametrize("coalesce", [True, False])
def test_eval_render_vs_reference_golden(golden, tag, n_fine, std, gain, coalesce):
    """Full-grid render, chunked (chunk_size_grid = 64*37 -> 9 chunks) and coalesced, vs the reference output."""
    g = golden("pipeline")
    B, H, W, n = 2, 16, 20, 48
    pipe = build_pipeline(H, W, n, n_fine, std, chunk=64 * 37).to(DEV)
    pipe.coalesce_chunks = coalesce
    load_synth_nets(pipe, seeds=(21, 22), gain=gain)
    poses, focal, image = syn.synth_camera(B, seed=5), torch.full((B, 1), 25.0), syn.synth_image(B, H, W, seed=4)
    with torch.no_grad():
        ev = pipe(poses=poses.to(DEV), focal_lengths=focal.to(DEV), image_rgb=image.to(DEV),
                  evaluation_mode=EvaluationMode.EVALUATION)
    assert ev["rendered_images"].shape == (B, H, W, 3)
    assert ev["rendered_depths"].shape == (B, H, W, 1) and ev["rendered_alpha_masks"].shape == (B, H, W, 1)
    err = float((ev["rendered_images"].cpu() - T(g[f"{tag}_eval_rendered_images"])).abs().max())
    derr = float((ev["rendered_depths"].cpu() - T(g[f"{tag}_eval_rendered_depths"])).abs().max())
    print(f"{tag} coalesce={coalesce}: rendered rgb max abs err {err:.2e}, depth {derr:.2e}")
    if gain == 1.0:
        # Rendered colour integrates the density along the ray: with random-init weights every one of the 128 / 192
        # samples is semi-transparent, and the rounding of the WEIGHTS to fp16 is a fixed smooth perturbation of the
        # density field, i.e. an error that is correlated along the ray and accumulates in the transmittance.  Measured:
        # 1.55e-3 (64 + 64 samples), 2.75e-3 (64 + 128 samples) on these fixtures; on trained weights (empty space +
        # surfaces) 3.8e-4 (test_trained_scale_render_vs_oracle) and 1.3e-4 on the 800x800 lego chunk
        # (test_full_size_render_properties), both asserted at the stated 2e-3.
        assert err <= (2e-3 if n_fine == 64 else 3e-3), err
        # depth = sum(w * z) with z in [2, 6]: the 16-bit-operand density error moves it by O(1e-2) relative
        assert derr <= 0.2, derr
        assert abs(float(ev["loss_rgb_mse"].mean().cpu()) - float(g[f"{tag}_eval_loss_rgb_mse"].mean())) <= 1e-3
    else:
        # "lego" tag = stress weights x3: raw densities reach O(1e3), the first sample with a positive density absorbs the
        # whole ray, and WHICH sample that is flips with the last bit of the pre-activation: individual pixels differ by
        # O(1) between ANY two arithmetics (fp32 vs fp64 of the reference itself included); only the mean is meaningful.
        # The lego.yml configuration at a meaningful scale is the "lego1" case above (max abs <= 2e-3).
        assert float((ev["rendered_images"].cpu() - T(g[f"{tag}_eval_rendered_images"])).abs().mean()) <= 2e-2
    torch.testing.assert_close(ev["rendered_alpha_masks"].cpu(), T(g[f"{tag}_eval_rendered_alpha_masks"]), rtol=0, atol=1e-6)
    assert set(ev) >= {"loss_rgb_huber", "loss_rgb_mse", "loss_prev_stage_rgb_huber", "loss_prev_stage_rgb_mse",
                       "rendered_images", "rendered_depths", "rendered_alpha_masks", "objective"}
    assert ev["objective"].shape == (B,)


def test_chunked_equals_coalesced_bitwise():
    B, H, W = 1, 24, 20
    pipe = build_pipeline(H, W, 32, 128, 0.0, chunk=64 * 50).to(DEV)
    load_synth_nets(pipe, seeds=(3, 4), gain=1.0)
    poses, focal = syn.synth_camera(B, seed=1).to(DEV), torch.full((B, 1), 30.0, device=DEV)
    outs = []
    for c in (True, False):
        pipe.coalesce_chunks = c
        with torch.no_grad():
            outs.append(pipe(poses=poses, focal_lengths=focal, evaluation_mode=EvaluationMode.EVALUATION))
    for k in ("rendered_images", "rendered_depths", "rendered_alpha_masks"):
        assert torch.equal(outs[0][k], outs[1][k]), k
    assert "loss_rgb_mse" not in outs[0]  # no image_rgb -> no losses (nerf_pipeline.py:249-282)


def test_zero_outputer_known_answer():
    """Reference tests/test_pipeline.py:128-151: zero density -> objective == 0 and render == background,
    with chunk_size_grid = 30 forcing the chunk loop, custom image size and background_density_bias 0."""
    from yanerf.pipelines import PIPELINES

    B, H, W = 2, 6, 10
    cfg = pipeline_cfg(H, W, 8, 16, 0.0, chunk=30)
    cfg.model = dict(type="ZeroOutputer")
    cfg.renderer.background_density_bias = 0.0
    cfg.renderer.bg_color = [0.0]
    pipe = PIPELINES.build(cfg).to(DEV)
    pipe.coalesce_chunks = False
    bg = syn.synth_image(B, H, W, seed=9).to(DEV)
    poses, focal = syn.synth_camera(B, seed=2).to(DEV), torch.full((B, 1), 12.0, device=DEV)
    for mode in (EvaluationMode.EVALUATION, EvaluationMode.TRAINING):
        preds = pipe(poses=poses, focal_lengths=focal, image_rgb=bg, bg_image_rgb=bg, evaluation_mode=mode)
        assert torch.equal(preds["objective"], torch.zeros(B, device=DEV))
        if mode == EvaluationMode.EVALUATION:
            assert torch.equal(preds["rendered_images"], bg)
    # custom image size at call time
    bg2 = syn.synth_image(B, 5, 7, seed=10).to(DEV)
    preds = pipe(poses=poses, focal_lengths=focal, image_rgb=bg2, bg_image_rgb=bg2, image_height=5, image_width=7,
                 evaluation_mode=EvaluationMode.EVALUATION)
    assert torch.equal(preds["rendered_images"], bg2)


def test_ray_sampler_shapes_and_ranges():
    """Reference tests/test_ray_sampler.py:27-99."""
    from yanerf.pipelines.ray_samplers import RAY_SAMPLERS

    B, H, W = 3, 12, 9
    rs = RAY_SAMPLERS.build(dict(type="RaySampler", image_width=W, image_height=H, n_pts_per_ray_training=6,
                                 n_pts_per_ray_evaluation=5, n_rays_per_image_sampled_from_mask=11, min_depth=1.0,
                                 max_depth=4.0))
    poses, focal = syn.synth_camera(B, seed=3).to(DEV), torch.full((B,), 10.0, device=DEV)
    tr = rs(poses, focal, EvaluationMode.TRAINING)
    assert tr.origins.shape == (B, 11, 1, 3) and tr.directions.shape == (B, 11, 1, 3)
    assert tr.lengths.shape == (B, 11, 1, 6) and tr.xys.shape == (B, 11, 1, 2)
    assert float(tr.lengths.min()) >= 1.0 and float(tr.lengths.max()) <= 4.0
    ev = rs(poses, focal, EvaluationMode.EVALUATION, min_depth=0.5, max_depth=2.0)
    assert ev.lengths.shape == (B, H, W, 5) and ev.xys.shape == (B, H, W, 2)
    assert float(ev.lengths.min()) >= 0.5 and float(ev.lengths.max()) <= 2.0
    ev2 = rs(poses, focal, EvaluationMode.EVALUATION, image_height=4, image_width=6)
    assert ev2.xys.shape == (B, 4, 6, 2)
    # xys -> flat index identity
    img = torch.rand(B, H, W, 3, device=DEV)
    assert torch.equal(sample_grid(img, ev.xys), img)
    got = sample_grid(img, tr.xys)
    assert torch.equal(sample_grid(scatter_rays_to_image(got, tr.xys, H, W), tr.xys), got)
    # per-layer sampling masks
    masks = torch.rand(B, 2, H, W, device=DEV)
    tr2 = rs(poses, focal, EvaluationMode.TRAINING, sampling_prob_mask=masks, n_rays_per_image=[4, 3])
    assert tr2.xys.shape == (B, 7, 1, 2)
    with pytest.raises(ValueError):
        rs(poses, focal, EvaluationMode.TRAINING, sampling_prob_mask=masks, n_rays_per_image=[4])
    # mask_crop `(B, 1, H, W)`: every picked pixel lies inside the mask, distinct while the mask has enough pixels (the
    # reference's own branch for this input raises, ray_sampler.py:82-97; the behaviour here is the evident intent)
    crop = torch.zeros(B, 1, H, W, device=DEV)
    crop[:, :, 2:9, 3:7] = 1.0  # 28 pixels >= 11 rays
    tr3 = rs(poses, focal, EvaluationMode.TRAINING, mask=crop)
    x, y = tr3.xys[..., 0].reshape(B, -1), tr3.xys[..., 1].reshape(B, -1)
    assert tr3.xys.shape == (B, 11, 1, 2)
    assert bool(((x >= 3) & (x < 7) & (y >= 2) & (y < 9)).all())
    flat = (y * W + x).long()
    assert all(len(set(row.tolist())) == 11 for row in flat)
    half = torch.nn.functional.interpolate(crop, size=[6, 5], mode="nearest")  # a mask at another resolution is resized
    tr4 = rs(poses, focal, EvaluationMode.TRAINING, mask=half)
    assert tr4.xys.shape == (B, 11, 1, 2)


def test_model_output_shapes():
    """Reference tests/test_models.py:19-55 (the latent_dim variant: tests/test_gpu_kernels.py::test_mlp_latent_codes...)."""
    from yanerf.pipelines.models import MODELS
    from tools.testing import LEGO_MLP

    mlp = MODELS.build(dict(LEGO_MLP)).to(DEV)
    o, d, z = torch.rand(2, 4, 5, 3, device=DEV), torch.rand(2, 4, 5, 3, device=DEV), torch.rand(2, 4, 5, 6, device=DEV)
    out = mlp(o, d, z)
    assert out["rays_densities"].shape == (2, 4, 5, 6, 1) and out["rays_features"].shape == (2, 4, 5, 6, 3)
    assert out["aux"] == {}
    with pytest.raises(ValueError):
        mlp(o, d, z, global_codes=torch.zeros(2, 1, 2, device=DEV))


def test_pipeline_global_codes():
    """Reference tests/test_pipeline.py:37-64 (`nerf_pipeline_cfg_with_conditional_mlp.py`): a latent_dim > 0 model, the
    IdentityMapper feature extractor hands `global_codes` through to the implicit functions, B = 3 images, chunk_size_grid
    30 forces the chunk loop in evaluation; training forward + backward reaches the code."""
    from yanerf.pipelines import PIPELINES
    from tools.testing import LEGO_MLP

    B, H, W = 3, 6, 10
    cfg = pipeline_cfg(H, W, 8, 16, 0.0, chunk=30)
    cfg.model = {**LEGO_MLP, "latent_dim": 4}
    cfg.feature_extractor = dict(type="IdentityMapper")
    cfg.renderer.blend_output = True
    cfg.ray_sampler.n_pts_per_ray_training = 16
    cfg.ray_sampler.n_pts_per_ray_evaluation = 16
    torch.manual_seed(0)
    pipe = PIPELINES.build(cfg).to(DEV)
    pipe.coalesce_chunks = False
    poses, focal = syn.synth_camera(B, seed=2).to(DEV), torch.full((B,), 12.0, device=DEV)
    image = syn.synth_image(B, H, W, seed=9).to(DEV)
    codes = torch.randn(B, 4, device=DEV, requires_grad=True)
    out = pipe(poses=poses, focal_lengths=focal, image_rgb=image, bg_image_rgb=image, evaluation_mode=EvaluationMode.TRAINING,
               global_codes=codes)
    assert out["objective"].shape == (B,)
    out["objective"].mean().backward()
    assert codes.grad is not None and torch.isfinite(codes.grad).all() and float(codes.grad.abs().sum()) > 0
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in pipe.parameters())
    with torch.no_grad():
        ev = pipe(poses=poses, focal_lengths=focal, image_rgb=image, evaluation_mode=EvaluationMode.EVALUATION,
                  global_codes=codes.detach())
        ev2 = pipe(poses=poses, focal_lengths=focal, image_rgb=image, evaluation_mode=EvaluationMode.EVALUATION,
                   global_codes=codes.detach().flip(0))
    assert ev["rendered_images"].shape == (B, H, W, 3) and torch.isfinite(ev["rendered_images"]).all()
    assert not torch.equal(ev["rendered_images"], ev2["rendered_images"])  # the code conditions the render
    with pytest.raises(ValueError, match="global codes"):
        pipe(poses=poses, focal_lengths=focal, evaluation_mode=EvaluationMode.EVALUATION)


def test_renderer_two_pass_same_model_runs():
    """Reference tests/test_renderer.py:17-57: same model object for both passes, per-ray bg colour."""
    from yanerf.pipelines.models import MODELS
    from yanerf.pipelines.renderers import RENDERERS
    from yanerf.pipelines.utils import PartialFunctionWrapper
    from tools.testing import LEGO_MLP

    mlp = PartialFunctionWrapper(MODELS.build(dict(LEGO_MLP))).to(DEV)
    ren = RENDERERS.build(dict(type="MultipassEmissionAbsorpsionRenderer", n_pts_per_ray_fine_training=4,
                               n_pts_per_ray_fine_evaluation=8, bg_color=[0.0, 0.0, 0.0]))
    o, d = torch.rand(3, 4, 5, 3, device=DEV), torch.rand(3, 4, 5, 3, device=DEV)
    z = torch.sort(torch.rand(3, 4, 5, 6, device=DEV) + 1.0, dim=-1)[0]
    xy, bg = torch.zeros(3, 4, 5, 2, device=DEV), torch.rand(3, 4, 5, 3, device=DEV)
    for mode, n in ((EvaluationMode.TRAINING, 10), (EvaluationMode.EVALUATION, 14)):
        out = ren(o, d, z, xy, bg, implicit_functions=[mlp, mlp], evaluation_mode=mode)
        assert out.features.shape == (3, 4, 5, 3) and out.depths.shape == (3, 4, 5, 1)
        assert out.aux["weights"].shape == (3, 4, 5, n) and out.prev_stage.aux["weights"].shape == (3, 4, 5, 6)
    with pytest.raises(ValueError, match="expects implicit functions"):
        ren(o, d, z, xy, bg, implicit_functions=[])


# --------------------------------------------------------------------------- training step
import contextlib


@contextlib.contextmanager
def inject_draws(multinomial, rand, randn_like):
    """Replay pre-generated draws into torch.multinomial / torch.rand / torch.randn_like in call order (the
    pipeline draws: pixels, stratified jitter, coarse noise, sample_pdf uniforms, fine noise)."""
    q = {"multinomial": list(multinomial), "rand": list(rand), "randn_like": list(randn_like)}
    saved = {k: getattr(torch, k) for k in q}

    def mk(name):
        def f(*a, **kw):
            return q[name].pop(0).clone()
        return f

    for k in q:
        setattr(torch, k, mk(k))
    try:
        yield
    finally:
        for k, v in saved.items():
            setattr(torch, k, v)
    for k, v in q.items():
        assert not v, f"unused injected draws for {k}"


@pytest.mark.parametrize("tag,n_fine,std,gain", [("fern", 64, 0.0, 1.0), ("lego1", 128, 0.2, 1.0), ("lego", 128, 0.2, 3.0)])
def test_train_forward_backward_vs_reference_golden(golden, tag, n_fine, std, gain):
    """TRAINING forward with the reference's draws replayed: losses vs the unmodified reference (golden), then
    backward: gradient summaries vs the reference's autograd."""
    g = golden("pipeline")
    B, H, W, n = 2, 16, 20, 48
    pipe = build_pipeline(H, W, n, n_fine, std, chunk=64 * 37).to(DEV)
    load_synth_nets(pipe, seeds=(21, 22), gain=gain)
    poses, focal, image = syn.synth_camera(B, seed=5), torch.full((B, 1), 25.0), syn.synth_image(B, H, W, seed=4)
    dr = {k: v.to(DEV) for k, v in syn.synth_draws(B, n, H * W, 64, n_fine, seed=6).items()}
    randn = [dr["noise0"], dr["noise1"]] if std > 0 else []
    with inject_draws([dr["pix"]], [dr["u_strat"], dr["u_pdf"]], randn):
        preds = pipe(poses=poses.to(DEV), focal_lengths=focal.to(DEV), image_rgb=image.to(DEV),
                     evaluation_mode=EvaluationMode.TRAINING)
    assert preds["objective"].shape == (B,)
    assert preds["rendered_images"].shape == (B, H, W, 3)
    tol = 2e-3 if gain == 1.0 else 5e-2
    for k in ("loss_rgb_mse", "loss_prev_stage_rgb_mse", "loss_rgb_huber", "objective"):
        got, ref = preds[k].detach().cpu(), T(g[f"{tag}_train_{k}"])
        print(tag, k, got.tolist(), ref.tolist())
        assert float((got - ref).abs().max()) <= tol, (k, got, ref)
    if gain == 1.0:
        err = float((preds["rendered_images"].cpu() - T(g[f"{tag}_train_rendered_images"])).abs().max())
        print(tag, "training forward (replayed draws): rendered rgb max abs err", err)
        assert err <= (2e-3 if n_fine == 64 else 3e-3), err  # same bound as the evaluation render: fp16 operands in both
    preds["objective"].mean().backward()
    # Every one of the 2 x 24 parameter tensors against the reference's autograd (golden summaries: sum, sum|.|, L2 norm;
    # three tensors in full).  The kernels run the layers with fp16 operands (11-bit significand) where the reference runs
    # fp32, so a small fraction of the ReLU units has the opposite sign (test_mlp_backward_vs_oracle_autograd quantifies it
    # per layer).  Bounds at xavier scale ("fern", "lego1"): per-tensor L2 norm and sum|.| within 2 %, the signed sum within
    # 2 % of sum|.| (measured worst: 0.96 %); the density head within 6 % (measured 4.3 % in the noise configuration: its
    # gradient sums d(loss)/d(sigma) over the samples whose noisy density is positive, and relu(sigma + noise) flips with the
    # 3e-4 error of sigma); full tensors cos >= 0.995 (measured >= 0.9971).  The x3 stress weights ("lego") are chaotic (see
    # the eval test): gradients there are only required to be finite.
    if gain != 1.0:
        for fn in pipe.implicit_functions:
            for name, p in fn._fn.named_parameters():
                assert p.grad is not None and torch.isfinite(p.grad).all(), name
        return
    k_norm, k_cos = 0.02, 0.995
    worst = dict(norm=0.0, l1=0.0, sum=0.0)
    failures = []
    for k, fn in enumerate(pipe.implicit_functions):
        for name, p in fn._fn.named_parameters():
            assert p.grad is not None and torch.isfinite(p.grad).all(), name
            gk = p.grad.detach().cpu().double().reshape(-1)
            ref = T(g[f"{tag}_grad{k}_{name}"]).double()
            rsum, rl1, rnorm = float(ref[0]), float(ref[1]), float(ref[2])
            if rnorm < 1e-12:
                assert float(gk.norm()) < 1e-9, name
                continue
            e_norm = abs(float(gk.norm()) / rnorm - 1)
            e_l1 = abs(float(gk.abs().sum()) / rl1 - 1)
            e_sum = abs(float(gk.sum()) - rsum) / rl1
            print(f"  {tag} net{k} {name:36s} norm {rnorm:9.3e} dev {e_norm:7.4f}  sum|.| dev {e_l1:7.4f}  sum dev/sum|.| {e_sum:7.4f}")
            worst = dict(norm=max(worst["norm"], e_norm), l1=max(worst["l1"], e_l1), sum=max(worst["sum"], e_sum))
            tol = 3 * k_norm if name.startswith("density_layer") else k_norm
            if not (e_norm <= tol and e_l1 <= tol and e_sum <= tol):
                failures.append((tag, k, name, round(e_norm, 4), round(e_l1, 4), round(e_sum, 4)))
        m = fn._fn
        for key, t in (("full_density_w", m.density_layer.weight), ("full_color2_w", m.color_layer[2].weight),
                       ("full_l0_b", m.xyz_encoder.mlp[0][0].bias)):
            gw, ref = t.grad.cpu().double().reshape(-1), T(g[f"{tag}_grad{k}_{key}"]).double().reshape(-1)
            cos = float((gw * ref).sum() / (gw.norm() * ref.norm()).clamp_min(1e-30))
            print(f"  {tag} net{k} {key}: cos {cos:.5f}, norm ratio {float(gw.norm() / ref.norm()):.4f}")
            if cos < k_cos:
                failures.append((tag, k, key, "cos", round(cos, 5)))
    print(tag, "worst relative deviations over all 48 gradient tensors:", worst)
    assert not failures, failures


def test_fused_trainer_converges_like_reference_runner():
    """Reference tests/test_runner.py:42-104: 2x2 image, 5+5 samples, Adam iterations -> objective < 0.01."""
    from yanerf.pipelines import PIPELINES
    from yanerf.runners import FusedTrainer

    torch.manual_seed(0)
    cfg = pipeline_cfg(2, 2, 4, 5, 0.0, chunk=0, min_depth=0.1, max_depth=2.0)
    cfg.ray_sampler.n_pts_per_ray_training = 5
    cfg.ray_sampler.n_pts_per_ray_evaluation = 5
    pipe = PIPELINES.build(cfg).to(DEV)
    trainer = FusedTrainer(pipe, lr=5e-3)
    batch = dict(poses=syn.synth_camera(1, seed=0).to(DEV), focal_lengths=torch.full((1, 1), 2.0, device=DEV),
                 image_rgb=torch.rand(1, 2, 2, 3, device=DEV))
    first = None
    from yanerf.runners.engine import reference_lr

    for it in range(200):
        # the runner's exponential schedule (runners/utils.py:65-109); a constant 5e-3 spikes now and then on this
        # 4-ray problem, whatever the arithmetic
        preds = trainer.train_step(batch, lr=reference_lr(it, init_lr=5e-3, min_lr=5e-4, lr_decay_rate=0.1, lr_decay_iters=200))
        if first is None:
            first = float(preds["objective"].detach().mean())
    with torch.no_grad():
        ev = pipe(**batch, evaluation_mode=EvaluationMode.EVALUATION)
    last = float(ev["objective"].mean())
    print("objective", first, "->", last)
    assert last < 0.01 and last < first


def test_fern_shape_llff_depth_tensors_and_custom_rays():
    """BASELINE configs[3]: fern.yml shape (378x504, 64+64 samples, 1024 rays) with per-image min/max depth TENSORS
    (reduced with .mean().item() like ray_sampler.py:280-283; there is no NDC warp in the reference)."""
    B, H, W = 2, 378, 504
    pipe = build_pipeline(H, W, 1024, 64, 0.0, 131072, min_depth=0.1, max_depth=8.0).to(DEV)
    load_synth_nets(pipe, seeds=(5, 6), gain=1.0)
    poses, focal = syn.synth_camera(B, seed=7).to(DEV), torch.full((B, 1), syn.FERN_FOCAL, device=DEV)
    near, far = torch.tensor([[1.2], [1.4]], device=DEV), torch.tensor([[12.0], [11.0]], device=DEV)
    image = syn.synth_image(B, H, W, seed=8).to(DEV)
    out = pipe(poses=poses, focal_lengths=focal, image_rgb=image, min_depth=near, max_depth=far,
               evaluation_mode=EvaluationMode.TRAINING)
    assert out["objective"].shape == (B,) and out["rendered_images"].shape == (B, H, W, 3)
    out["objective"].mean().backward()
    assert all(torch.isfinite(p.grad).all() for p in pipe.parameters())
    with torch.no_grad():
        ev = pipe(poses=poses[:1], focal_lengths=focal[:1], image_rgb=image[:1], min_depth=near[:1], max_depth=far[:1],
                  evaluation_mode=EvaluationMode.EVALUATION)
    assert ev["rendered_images"].shape == (1, H, W, 3) and torch.isfinite(ev["rendered_images"]).all()
    d = ev["rendered_depths"]
    assert float(d.min()) >= 1.2 - 1e-3 and float(d.max()) <= 12.0 + 1e-3
    # oracle on a slab of rays of the same image (chunk plan of fern: 94 x 2027 rays)
    spec = oracle_spec(H, W, 64, 0.0, 131072, min_depth=1.2, max_depth=12.0)
    nets = [{k: v.detach().cpu() for k, v in fn._fn.state_dict().items()} for fn in pipe.implicit_functions]
    s, e = 90000, 90000 + 2027
    ref = O.render_image(nets, spec, poses[:1].cpu(), focal[:1].cpu(), ray_slice=(s, e))
    got = ev["rendered_images"].reshape(1, H * W, 3)[:, s:e].cpu()
    assert float((got - ref["features"]).abs().max()) <= 2e-3


def test_full_size_render_properties():
    """BASELINE configs[1] at full size (800x800, 64+128): size-independent properties instead of an oracle run.
    determinism (bitwise), sorted refined depths, weights in [0,1] summing to opacity, linearity of the colour
    compositing in rgb, chunk-invariance on a row band."""
    from yanerf import ops

    H = W = 800
    pipe = build_pipeline(H, W, 4096, 128, 0.2, 131072).to(DEV)
    load_synth_nets(pipe, seeds=(0, 1), gain=1.0)
    poses, focal = syn.synth_camera(1, seed=0, jitter=0.0).to(DEV), torch.full((1, 1), syn.LEGO_FOCAL, device=DEV)
    with torch.no_grad():
        a = pipe(poses=poses, focal_lengths=focal, evaluation_mode=EvaluationMode.EVALUATION)
        b = pipe(poses=poses, focal_lengths=focal, evaluation_mode=EvaluationMode.EVALUATION)
    for k in ("rendered_images", "rendered_depths", "rendered_alpha_masks"):
        assert torch.equal(a[k], b[k]), f"{k} not deterministic"
        assert torch.isfinite(a[k]).all()
    assert a["rendered_images"].shape == (1, H, W, 3)
    assert float(a["rendered_alpha_masks"].min()) == 1.0  # background_opacity 1e10 saturates every ray
    assert float(a["rendered_depths"].min()) >= 2.0 - 1e-4 and float(a["rendered_depths"].max()) <= 6.0 + 1e-4
    # renderer-level properties on one band of rays
    bundle = pipe.ray_sampler(poses, focal, EvaluationMode.EVALUATION)
    sl = slice(300, 310)
    o, d, z, xy = (t[:, sl].contiguous() for t in bundle)
    with torch.no_grad():
        out = pipe.renderer(o, d, z, xy, None, implicit_functions=pipe.implicit_functions,
                            evaluation_mode=EvaluationMode.EVALUATION)
    assert torch.equal(out.features, a["rendered_images"][:, sl]), "band render differs from the full-image render"
    w = out.aux["weights"]
    assert w.shape[-1] == 192 and float(w.min()) >= 0.0
    torch.testing.assert_close(w.sum(-1, keepdim=True), out.alpha_masks, rtol=0, atol=2e-5)
    zf = ops.sample_pdf_merge(z.reshape(-1, 64), out.prev_stage.aux["weights"].reshape(-1, 64), 128, None)[0]
    assert bool((zf[:, 1:] >= zf[:, :-1]).all()), "refined depths not ascending"
    # linearity of compositing in the colours
    R = zf.shape[0]
    sig, c1, c2 = torch.randn(R, 192, device=DEV), torch.rand(R, 192, 3, device=DEV), torch.rand(R, 192, 3, device=DEV)
    cfg = ops.march_cfg(1e10, 1e-6, 0.0, False, False, (0.0, 0.0, 0.0))
    dd = d.reshape(-1, 3)
    f1, f2, f12 = (ops.composite(sig, c, zf, dd, cfg)[0] for c in (c1, c2, 2 * c1 + 3 * c2))
    torch.testing.assert_close(f12, 2 * f1 + 3 * f2, rtol=1e-5, atol=1e-5)
    # one reference-sized chunk (2045 rays: chunk 150 of the 313 the reference would loop over) against the CPU oracle
    nets = [{k: v.detach().cpu() for k, v in fn._fn.state_dict().items()} for fn in pipe.implicit_functions]
    n_chunks, per = O.chunk_plan(H * W, 64, 131072)
    assert (n_chunks, per) == (313, 2045)
    s0 = 150 * per
    ref = O.render_image(nets, oracle_spec(H, W, 128, 0.2, 131072), poses.cpu(), focal.cpu(), ray_slice=(s0, s0 + per))
    got = a["rendered_images"].reshape(1, H * W, 3)[:, s0:s0 + per].cpu()
    err = float((got - ref["features"]).abs().max())
    print(f"800x800 lego render, chunk 150 ({per} rays) vs oracle: max abs rgb err {err:.2e}")
    assert err <= 2e-3, err


def test_trained_scale_render_vs_oracle():
    """SURVEY 7-C / VERDICT r1: the 2e-3 bound on TRAINED weights, not only on xavier-scale ones.  The lego.yml
    architecture (64 + 128 samples, density noise 0.2) is trained for 400 fused steps on an analytic 8-view scene (a shaded
    sphere), then one whole 48 x 48 view (2304 rays > one reference chunk of 2045) is rendered by the kernels (fp16
    operands, the evaluation default) and by the CPU oracle (fp32) from the SAME trained weights."""
    from yanerf.runners import FusedTrainer

    torch.manual_seed(3)
    H = W = 48
    n_views, focal = 8, 60.0
    pipe = build_pipeline(H, W, 1024, 128, 0.2, 131072).to(DEV)
    poses = syn.orbit_cameras(n_views)
    images = syn.sphere_scene_images(poses, focal, H, W)
    trainer = FusedTrainer(pipe, lr=5e-4)
    fl = torch.full((1, 1), focal, device=DEV)
    for it in range(400):
        v = it % n_views
        trainer.train_step(dict(poses=poses[v:v + 1].to(DEV), focal_lengths=fl, image_rgb=images[v:v + 1].to(DEV)))
    trainer.finish()
    nets = [{k: v.detach().cpu().clone() for k, v in fn._fn.state_dict().items()} for fn in pipe.implicit_functions]
    scale = max(float(sd["density_layer.weight"].abs().max()) for sd in nets)
    with torch.no_grad():
        ev = pipe(poses=poses[:1].to(DEV), focal_lengths=fl, image_rgb=images[:1].to(DEV), evaluation_mode=EvaluationMode.EVALUATION)
        ref = O.render_image(nets, oracle_spec(H, W, 128, 0.2, 131072), poses[:1], torch.full((1, 1), focal))
    got = ev["rendered_images"].reshape(1, H * W, 3).cpu()
    err = (got - ref["features"]).abs()
    mse = float(((got - ref["features"]) ** 2).mean())
    psnr_vs_oracle = -10.0 * np.log10(max(mse, 1e-20))
    psnr_vs_gt = -10.0 * np.log10(float(((ref["features"] - images[:1].reshape(1, H * W, 3)) ** 2).mean()))
    print(f"trained-scale: oracle-vs-truth PSNR {psnr_vs_gt:.1f} dB; kernel-vs-oracle max abs {float(err.max()):.2e}, "
          f"mean {float(err.mean()):.2e}, PSNR {psnr_vs_oracle:.1f} dB; max |w_density| {scale:.2f}")
    assert psnr_vs_gt > 14.0, "the training did not move the weights away from their initial scale"
    assert float(err.max()) <= 2e-3, float(err.max())
    assert psnr_vs_oracle >= 55.0, psnr_vs_oracle
    derr = float((ev["rendered_depths"].reshape(1, H * W, 1).cpu() - ref["depths"]).abs().max())
    assert derr <= 0.05, derr


def test_training_fits_a_synthetic_image():
    """End-to-end optimisation with the fused step (kernels fwd + bwd + Adam): a lego.yml-architecture pipeline on a
    24x24 two-tone image with one fixed camera must reach PSNR > 22 dB within 400 steps (it memorises the view)."""
    from yanerf.pipelines import PIPELINES
    from yanerf.runners import FusedTrainer
    from yanerf.runners.apis import create_stats

    torch.manual_seed(1)
    H = W = 24
    cfg = pipeline_cfg(H, W, 256, 32, 0.0, chunk=131072)
    cfg.ray_sampler.n_pts_per_ray_training = 32
    cfg.ray_sampler.n_pts_per_ray_evaluation = 32
    pipe = PIPELINES.build(cfg).to(DEV)
    trainer = FusedTrainer(pipe, lr=5e-4)
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    img = torch.where(((xx - 12) ** 2 + (yy - 12) ** 2 < 49)[..., None], torch.tensor([0.9, 0.2, 0.1]), torch.tensor([0.1, 0.3, 0.8]))
    batch = dict(poses=syn.synth_camera(1, seed=0, jitter=0.0).to(DEV), focal_lengths=torch.full((1, 1), 30.0, device=DEV),
                 image_rgb=img[None].to(DEV))
    first = None
    for it in range(400):
        preds = trainer.train_step(batch)
        if it == 0:
            first = create_stats(preds)["loss_rgb_psnr"]
    trainer.finish()
    with torch.no_grad():
        ev = pipe(**batch, evaluation_mode=EvaluationMode.EVALUATION)
    psnr = create_stats(ev)["loss_rgb_psnr"]
    print(f"PSNR {first:.2f} -> {psnr:.2f} dB after 400 fused steps")
    assert psnr > 22.0 and psnr > first + 8.0


def test_cuda_graph_training_matches_eager_and_converges():
    """The captured-graph step (fused ray kernel, forward, backward, Adam with device-side step / lr, weight re-pack) against
    the eager step ON IDENTICAL DRAWS: both trainers key their in-kernel Philox streams with the same seed, so pixels,
    jitter, density noise and inverse-CDF uniforms coincide step by step.  The two loss trajectories must then agree (the
    only difference left is the order of the fp32 atomics in the weight-gradient flushes), both must reach PSNR > 22 dB on
    the 24x24 image, and lr = 0 (a host value the graph reads from device memory) must freeze the weights."""
    from yanerf.pipelines import PIPELINES
    from yanerf.runners import FusedTrainer
    from yanerf.runners.apis import create_stats

    H = W = 24
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    img = torch.where(((xx - 12) ** 2 + (yy - 12) ** 2 < 49)[..., None], torch.tensor([0.9, 0.2, 0.1]), torch.tensor([0.1, 0.3, 0.8]))
    results = {}
    for graph in (False, True):
        torch.manual_seed(1)
        cfg = pipeline_cfg(H, W, 256, 32, 0.1, chunk=131072)
        cfg.ray_sampler.n_pts_per_ray_training = 32
        cfg.ray_sampler.n_pts_per_ray_evaluation = 32
        pipe = PIPELINES.build(cfg).to(DEV)
        trainer = FusedTrainer(pipe, lr=5e-4, use_cuda_graph=graph)
        assert pipe.ray_sampler.fused_pixel_sampler and trainer.rng is not None
        trainer.rng.seed(777)
        batch = dict(poses=syn.synth_camera(1, seed=0, jitter=0.0).to(DEV), focal_lengths=torch.full((1, 1), 30.0, device=DEV),
                     image_rgb=img[None].to(DEV))
        losses = []
        for it in range(300):
            preds = trainer.train_step(batch, lr=5e-4 if it < 200 else 0.0)
            if it == 250:
                frozen = trainer.flat.clone()
            if it < 40 or it % 50 == 49 or it >= 298:
                losses.append(float(preds["objective"].detach().mean()))
        trainer.finish()
        with torch.no_grad():
            ev = pipe(**batch, evaluation_mode=EvaluationMode.EVALUATION)
        results[graph] = (losses, create_stats(ev)["loss_rgb_psnr"], trainer.flat.clone())
        assert trainer.step_count == 300 and int(trainer.rng.state[1]) == 300
        assert torch.equal(frozen, trainer.flat)  # lr = 0 freezes the weights
    (le, pe, fe), (lg, pg, fg) = results[False], results[True]
    print("eager", [round(x, 5) for x in le[:6]], "...", pe, "graph", [round(x, 5) for x in lg[:6]], "...", pg)
    assert le[0] == lg[0] or abs(le[0] - lg[0]) <= 1e-6 * abs(le[0])  # step 0: same weights, same draws
    for k, (a, b) in enumerate(zip(le[:40], lg[:40])):  # the first 40 steps track each other closely
        assert abs(a - b) <= 0.03 * max(abs(a), 1e-3) + 2e-4, (k, a, b)
    assert pe > 22.0 and pg > 22.0  # (after 300 chaotic steps the two runs differ by a few dB, like any two runs)


def test_ray_slab_sharded_render_equals_unsharded(monkeypatch):
    """rendering one image as 3 ray slabs (one per emulated rank) and stitching the slabs gives the
    same image, bit for bit, as the unsharded render: rays are independent and slabs are whole MLP tiles."""
    import yanerf.pipelines.nerf_pipeline as NP

    H, W = 50, 44  # 2200 rays -> slabs of 768, 768, 664
    pipe = build_pipeline(H, W, 256, 64, 0.0, 4096).to(DEV)
    load_synth_nets(pipe, seeds=(3, 4), gain=1.0)
    batch = dict(poses=syn.synth_camera(1, seed=2).to(DEV), focal_lengths=torch.full((1, 1), 60.0, device=DEV),
                 image_rgb=syn.synth_image(1, H, W, seed=3).to(DEV))
    with torch.no_grad():
        ref = pipe(**batch, evaluation_mode=EvaluationMode.EVALUATION)
        acc = {}

        def fake_gather(local, n_rays, per, group=None):
            rank = pipe.ray_shard[0]
            full = local.new_zeros(local.shape[0], n_rays, local.shape[2])
            s = min(rank * per, n_rays)
            full[:, s:s + local.shape[1]] = local
            return full

        monkeypatch.setattr(NP, "gather_slabs", fake_gather)
        for rank in range(3):
            pipe.ray_shard = (rank, 3, None)
            out = pipe(**batch, evaluation_mode=EvaluationMode.EVALUATION)
            # without a process group the all-reduce is the identity: every emulated rank reports its slab's share
            for k in ("rendered_images", "rendered_depths", "rendered_alpha_masks", "loss_rgb_mse", "loss_prev_stage_rgb_mse"):
                acc[k] = out[k] if k not in acc else acc[k] + out[k]
        pipe.ray_shard = None
    for k, v in acc.items():
        if k.startswith("rendered"):
            assert torch.equal(v, ref[k]), k
        else:  # slab sums of squared errors / n_rays add up to the full-image mean (fp32 summation order aside)
            torch.testing.assert_close(v, ref[k], rtol=1e-5, atol=1e-8)


def test_checkpoint_resume_continues_the_same_trajectory():
    """save in the reference's checkpoint layout after 4 steps, load into a fresh pipeline + trainer,
    continue 3 steps: same weights as the uninterrupted run (up to the fp32 atomics of the gradient reduction)."""
    import io

    from yanerf.pipelines import PIPELINES
    from yanerf.runners import FusedTrainer

    H = W = 16
    batch = dict(poses=syn.synth_camera(1, seed=0, jitter=0.0).to(DEV), focal_lengths=torch.full((1, 1), 20.0, device=DEV),
                 image_rgb=syn.synth_image(1, H, W, seed=5).to(DEV))

    def fresh():
        torch.manual_seed(11)
        cfg = pipeline_cfg(H, W, 128, 32, 0.0, chunk=131072)
        cfg.ray_sampler.n_pts_per_ray_training = 32
        pipe = PIPELINES.build(cfg).to(DEV)
        return pipe, FusedTrainer(pipe, lr=1e-3)

    def reseed(trainer):
        trainer.rng.seed(99, step=trainer.step_count)  # both runs continue on the same in-kernel draw stream

    pipe_a, tr_a = fresh()
    for _ in range(4):
        tr_a.train_step(batch)
    buf = io.BytesIO()
    torch.save(tr_a.state_dict(epoch=0), buf)  # the file scripts/run.py:416-422 writes
    reseed(tr_a)
    for _ in range(3):
        tr_a.train_step(batch)
    tr_a.finish()

    pipe_b, tr_b = fresh()
    buf.seek(0)
    assert tr_b.load_state_dict(torch.load(buf, map_location="cpu")) == 1
    assert tr_b.step_count == 4
    reseed(tr_b)
    for _ in range(3):
        tr_b.train_step(batch)
    tr_b.finish()
    assert tr_b.step_count == 7
    # fp32 atomics in the gradient reductions are order-dependent, and Adam turns a tiny gradient of either sign into a
    # full-size step, so a handful of weights differ by up to a step between ANY two runs; the trajectories agree in norm
    init = fresh()[1].flat
    diff = float((tr_a.flat - tr_b.flat).norm())
    moved = float((tr_a.flat - init).norm())
    print("resume: |W_a - W_b| =", diff, " distance travelled |W_a - W_0| =", moved)
    assert diff <= 0.05 * moved


def test_runner_epoch_loop_on_device_feed():
    """Runner loop end to end (runners/apis.py train_one_epoch / eval_one_epoch) fed by the device-resident scene
    cache, graph-captured step: the loss of a 6-view synthetic scene falls and the eval stats come out per key."""
    from yanerf.pipelines import PIPELINES
    from yanerf.runners import DeviceSceneFeed, FusedTrainer
    from yanerf.runners.apis import eval_one_epoch, train_one_epoch

    torch.manual_seed(0)
    H = W = 20
    n = 6
    cfg = pipeline_cfg(H, W, 200, 32, 0.0, chunk=131072)
    cfg.ray_sampler.n_pts_per_ray_training = 32
    cfg.ray_sampler.n_pts_per_ray_evaluation = 32
    pipe = PIPELINES.build(cfg).to(DEV)
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    img = torch.where(((xx - 10) ** 2 + (yy - 10) ** 2 < 36)[..., None], torch.tensor([0.8, 0.7, 0.1]), torch.tensor([0.2, 0.2, 0.6]))
    feed = DeviceSceneFeed(syn.synth_camera(1, seed=0, jitter=0.0).expand(n, -1, -1), 25.0, img[None].expand(n, -1, -1, -1),
                           DEV, shuffle=True, seed=1)
    trainer = FusedTrainer(pipe, lr=5e-4, use_cuda_graph=True)
    conf = dict(init_lr=1e-3, min_lr=1e-4, lr_decay_type="exponential", lr_decay_rate=0.1, lr_decay_iters=120, num_iters=120,
                warmup_steps=0, warmup_lr=0.0, linear_scale=True)
    first = train_one_epoch(trainer, feed, conf, epoch=0, iters_per_epoch=len(feed))
    for epoch in range(1, 20):
        feed.set_epoch(epoch)
        last = train_one_epoch(trainer, feed, conf, epoch=epoch, iters_per_epoch=len(feed))
    trainer.finish()
    stats = eval_one_epoch(pipe, [feed.batch(0), feed.batch(1)], dataset_len=2)
    print("epoch stats", first["objective"], "->", last["objective"], "eval psnr", stats["loss_rgb_psnr"])
    assert trainer.step_count == 120 and last["objective"] < 0.25 * first["objective"]
    assert stats["loss_rgb_psnr"] > 18.0 and "loss_prev_stage_rgb_mse" in stats


@pytest.mark.parametrize("case", ["bgdepth", "hardbg", "mask", "custom"])
def test_pipeline_optional_inputs_vs_reference_golden(golden, case):
    """The pipeline's optional inputs and renderer switches against the UNMODIFIED reference
    (tests/golden/make_golden_variants.py): per-pixel background image + depth map with blend_output, hard white
    background, `sampling_prob_mask` training picks, custom evaluation grid / depth range.  Same weights, inputs and
    replayed draws; rendered values within the 16-bit-operand bound of the main golden tests, losses within 2e-3,
    gradient norms within 2 % (density head 6 %)."""
    g = golden("pipeline_variants")
    B, H, W, n, n_fine = 2, 16, 20, 48, 64
    over = {"bgdepth": dict(blend_output=True), "hardbg": dict(hard_background=True, bg_color=[1.0, 1.0, 1.0])}.get(case)
    pipe = build_pipeline(H, W, n, n_fine, 0.0, chunk=64 * 37, renderer=over).to(DEV)
    load_synth_nets(pipe, seeds=(31, 32), gain=1.0)
    batch = dict(poses=syn.synth_camera(B, seed=8).to(DEV), focal_lengths=torch.full((B, 1), 25.0, device=DEV),
                 image_rgb=syn.synth_image(B, H, W, seed=9).to(DEV))
    if case == "bgdepth":
        batch.update(bg_image_rgb=syn.synth_image(B, H, W, seed=10).to(DEV),
                     depth_map=(2.0 + 4.0 * syn.synth_image(B, H, W, seed=11)[..., :1]).to(DEV))
    dr = {k: v.to(DEV) for k, v in syn.synth_draws(B, n, H * W, 64, n_fine, seed=12).items()}

    def compare(tag, preds, keys_tol):
        for k, tol in keys_tol.items():
            ref = T(g[f"{case}_{tag}_{k}"])
            got = preds[k].detach().cpu()
            assert got.shape == ref.shape, (k, got.shape, ref.shape)
            err, mean = float((got - ref).abs().max()), float((got - ref).abs().mean())
            print(f"{case} {tag} {k}: max abs err {err:.3e}, mean {mean:.3e} (ref magnitude {float(ref.abs().max()):.3e})")
            assert err <= tol and mean <= tol / 4, (case, tag, k, err, mean)

    loss_tol = {k: 2e-3 for k in ("loss_rgb_mse", "loss_prev_stage_rgb_mse", "loss_rgb_huber", "loss_prev_stage_rgb_huber",
                                  "objective")}
    image_tol = dict(rendered_images=3e-3, rendered_alpha_masks=3e-3, rendered_depths=5e-2)  # depths in [2, 6]: <= 1 %
    if case == "hardbg":
        # the last sample absorbs what is left of the ray and shows the WHITE background: the transmittance error of the
        # semi-transparent random-init field (see test_eval_render_vs_reference_golden) is multiplied by 1.0 instead of a
        # dim colour.  Measured 6.3e-3 max, mean below 1e-3.
        image_tol["rendered_images"] = 1e-2
    if case in ("bgdepth", "hardbg", "mask"):
        extra = dict(sampling_prob_mask=T(g["mask_prob"]).to(DEV)) if case == "mask" else {}
        with inject_draws([dr["pix"]], [dr["u_strat"], dr["u_pdf"]], []):
            preds = pipe(**batch, **extra, evaluation_mode=EvaluationMode.TRAINING)
        tol = dict(loss_tol, **image_tol)
        if case == "bgdepth":
            tol.update(loss_depth_abs=5e-2, loss_prev_stage_depth_abs=5e-2)
        compare("train", preds, tol)
        if case != "mask":
            preds["objective"].mean().backward()
            bad = []
            for k, fn in enumerate(pipe.implicit_functions):
                for name, p in fn._fn.named_parameters():
                    ref = T(g[f"{case}_grad{k}_{name}"]).double()
                    gk = p.grad.detach().cpu().double().reshape(-1)
                    if float(ref[2]) < 1e-12:
                        assert float(gk.norm()) < 1e-9, name
                        continue
                    dev = abs(float(gk.norm()) / float(ref[2]) - 1)
                    if dev > (0.06 if name.startswith("density_layer") else 0.02):
                        bad.append((k, name, round(dev, 4)))
            assert not bad, bad
    if case in ("bgdepth", "hardbg"):
        with torch.no_grad():
            ev = pipe(**batch, evaluation_mode=EvaluationMode.EVALUATION)
        tol = dict(image_tol, loss_rgb_mse=2e-3, loss_prev_stage_rgb_mse=2e-3, objective=2e-3)
        if case == "bgdepth":
            tol.update(loss_depth_abs=5e-2, loss_prev_stage_depth_abs=5e-2)
        compare("eval", ev, tol)
    if case == "custom":
        with torch.no_grad():
            ev = pipe(**batch, image_height=12, image_width=10, min_depth=1.5, max_depth=5.0,
                      evaluation_mode=EvaluationMode.EVALUATION)
        assert ev["rendered_images"].shape == (B, 12, 10, 3)
        compare("eval", ev, dict(image_tol, loss_rgb_mse=2e-3, loss_prev_stage_rgb_mse=2e-3, objective=2e-3))
