"""GPU parity tests (run on the B200 box): every kernel is called through the C ABI (via yanerf.ops) and
compared with the CPU oracle / the golden fixtures produced by the unmodified reference.

Tolerances (BASELINE.json north_star):
  * sample_pdf / refiner: indices AND samples bit-exact;
  * compositing: rtol 1e-5 + atol 1e-7 in fp32 (SURVEY §7-D explains the absolute floor on `weights`);
  * MLP outputs / rendered RGB: max abs <= 2e-3 (16-bit tensor-core operands, fp32 accumulate).
"""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O
from tools import synthetic as syn

pytestmark = pytest.mark.gpu

DEV = "cuda"


def T(a):
    return torch.from_numpy(np.asarray(a))


def close(got, ref, rtol, atol, what=""):
    got, ref = got.detach().cpu().double(), torch.as_tensor(ref).double()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    bad = (got - ref).abs() > atol + rtol * ref.abs()
    assert not bad.any(), f"{what}: {int(bad.sum())}/{bad.numel()} off, max abs {float((got - ref).abs().max()):.3e}"


def same(got, ref, what=""):
    got, ref = got.detach().cpu(), torch.as_tensor(ref)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    n = int((got != ref).sum())
    assert n == 0, f"{what}: {n}/{got.numel()} differ, max abs {float((got.double() - ref.double()).abs().max()):.3e}"


# --------------------------------------------------------------------------- sample_pdf
@pytest.mark.parametrize("tag,n", [("lego", 128), ("fern", 64), ("wide", 128)])
def test_sample_pdf_golden(golden, tag, n):
    from yanerf import ops

    g = golden("sample_pdf")
    bins, w, u = (T(g[f"{tag}_{k}"]).to(DEV) for k in ("bins", "weights", "u"))
    s, inds, flag = ops.sample_pdf(bins, w, n, None, want_inds=True)
    same(inds, g[f"{tag}_inds_det"], "det inds")
    same(s, g[f"{tag}_det"], "det samples")
    s, inds, flag = ops.sample_pdf(bins, w, n, u, want_inds=True)
    same(inds, g[f"{tag}_inds_rnd"], "rnd inds")
    same(s, g[f"{tag}_rnd"], "rnd samples")
    assert int(flag.item()) == 0


@pytest.mark.parametrize("tag,n", [("lego", 128), ("fern", 64)])
def test_refiner_golden(golden, tag, n):
    from yanerf import ops

    g = golden("refiner")
    z, w, u = T(g[f"{tag}_z"]).reshape(-1, 64).to(DEV), T(g[f"{tag}_w"]).reshape(-1, 64).to(DEV), T(g[f"{tag}_u"]).to(DEV)
    R = z.shape[0]
    same(ops.sample_pdf_merge(z, w, n, None)[0], g[f"{tag}_det"].reshape(R, -1), "det")
    same(ops.sample_pdf_merge(z, w, n, u)[0], g[f"{tag}_rnd"].reshape(R, -1), "rnd")
    same(ops.sample_pdf_merge(z, w, n, None, add_input_samples=False)[0], g[f"{tag}_noadd"].reshape(R, -1), "noadd")


@pytest.mark.parametrize("P,n,R", [(64, 128, 20000), (192, 128, 4000), (64, 64, 5000), (7, 5, 300), (3, 1, 50)])
def test_refiner_vs_oracle_random(P, n, R):
    """Index + sample bit-exactness on seeded random rays (sparse weights, empty rays, repeated depths)."""
    from yanerf import ops

    rs = np.random.RandomState(P * 1000 + n)
    z = np.sort(2 + 4 * rs.uniform(size=(R, P)).astype(np.float32), axis=-1)
    z[: R // 10, P // 2:] = z[: R // 10, P // 2: P // 2 + 1]  # ties
    w = rs.uniform(size=(R, P)).astype(np.float32) ** 6
    w[rs.uniform(size=w.shape) < 0.4] = 0.0
    w[0] = 0.0
    u = np.minimum(rs.uniform(size=(R, n)).astype(np.float32), np.float32(1 - 2 ** -24))
    zt, wt, ut = T(z), T(w), T(u)
    for uu in (None, ut):
        ref, ref_inds = O.refine_lengths(zt, wt, n, uu)
        got, inds, flag = ops.sample_pdf_merge(zt.to(DEV), wt.to(DEV), n, None if uu is None else uu.to(DEV), want_inds=True)
        same(inds, ref_inds, f"inds P={P} n={n} det={uu is None}")
        same(got, ref, f"lengths P={P} n={n} det={uu is None}")
        assert int(flag.item()) == 0


def test_refiner_unsorted_input_and_negative_weights():
    from yanerf import ops
    from yanerf.pipelines.renderers.utils import RayPointRefiner

    rs = np.random.RandomState(5)
    z = T((2 + 4 * rs.uniform(size=(64, 32))).astype(np.float32))  # NOT sorted
    w = T(rs.uniform(size=(64, 32)).astype(np.float32))
    ref, _ = O.refine_lengths(z, w, 16, None)
    same(ops.sample_pdf_merge(z.to(DEV), w.to(DEV), 16, None)[0], ref, "unsorted input")
    w[3, 5] = -1.0
    ref_raises = False
    try:
        O.refine_lengths(z, w, 16, None)
    except ValueError:
        ref_raises = True
    assert ref_raises
    zz = z.to(DEV).reshape(1, 64, 1, 32)
    with pytest.raises(ValueError, match="Negative weights"):
        RayPointRefiner(16, False)(zz, zz, zz, zz, w.to(DEV).reshape(1, 64, 1, 32))


# --------------------------------------------------------------------------- compositing
VARIANTS = {
    "lego": O.RaymarcherSpec(background_density_bias=1e-6, bg_color=(0.0, 0.0, 0.0)),
    "blend": O.RaymarcherSpec(blend_output=True, bg_color=(0.0, 0.0, 0.0)),
    "hard": O.RaymarcherSpec(hard_background=True, bg_color=(0.3, 0.5, 0.7)),
}


def _cfg(spec, std):
    from yanerf import ops

    return ops.march_cfg(spec.background_opacity, spec.background_density_bias, std, spec.blend_output,
                         spec.hard_background, spec.bg_color)


@pytest.mark.parametrize("P", [64, 192])
@pytest.mark.parametrize("name", list(VARIANTS))
@pytest.mark.parametrize("with_noise", [0, 1])
@pytest.mark.parametrize("with_bg", [0, 1])
def test_composite_golden(golden, P, name, with_noise, with_bg):
    from yanerf import ops

    g = golden("raymarch")
    sig, rgb, z, d, noise, bg = (T(g[f"P{P}_{k}"]).to(DEV) for k in ("sigma", "rgb", "z", "d", "noise", "bg"))
    std = 0.2 if with_noise else 0.0
    f, dep, op, w = ops.composite(sig[..., 0].contiguous(), rgb, z, d, _cfg(VARIANTS[name], std),
                                  noise if with_noise else None, bg if with_bg else None)
    key = f"P{P}_{name}_n{with_noise}_b{with_bg}"
    # the reference's own fp32-vs-fp64 self-consistency for these sums is ~1e-6: each of the
    # P terms carries the 2^-24 quantisation of T = 1 - (1 - E)
    close(f, g[key + "_feat"], 1e-5, 4e-6, "features")
    close(dep, g[key + "_depth"], 1e-5, 4e-6, "depths")
    close(op, g[key + "_opac"], 1e-5, 1e-7, "opacities")
    close(w, g[key + "_w"], 1e-5, 1e-7, "weights")


@pytest.mark.parametrize("P", [64, 192, 37])
@pytest.mark.parametrize("name", list(VARIANTS))
@pytest.mark.parametrize("with_bg", [0, 1])
def test_composite_backward_vs_oracle_autograd(P, name, with_bg):
    """Analytic backward vs torch autograd through the oracle, all four outputs receiving gradient."""
    from yanerf import ops

    rs = np.random.RandomState(100 + P)
    R = 257
    sig = T((3 * rs.standard_normal(size=(R, P)) + 0.5).astype(np.float32))
    rgb = T(rs.uniform(size=(R, P, 3)).astype(np.float32))
    z = T(np.sort(2 + 4 * rs.uniform(size=(R, P)), axis=-1).astype(np.float32))
    d = T(rs.standard_normal(size=(R, 3)).astype(np.float32))
    noise = T(rs.standard_normal(size=(R, P)).astype(np.float32))
    bg = T(rs.uniform(size=(R, 3)).astype(np.float32))
    gf, gd, go, gw = (T(rs.standard_normal(size=s).astype(np.float32)) for s in ((R, 3), (R, 1), (R, 1), (R, P)))
    spec = VARIANTS[name]
    # thin densities so that opacity does not saturate and every gradient path is exercised
    scale = 0.02 if name != "lego" else 1.0
    a = (sig * scale).clone().requires_grad_(True)
    b = rgb.clone().requires_grad_(True)
    spec_o = O.RaymarcherSpec(background_opacity=1e10 if name == "lego" else 3.0, background_density_bias=spec.background_density_bias,
                              blend_output=spec.blend_output, hard_background=spec.hard_background, bg_color=spec.bg_color)
    f, dep, op, w = O.raymarch(a, b, z, d, spec_o, noise, 0.2, bg if with_bg else None)
    (f * gf).sum().add((dep * gd).sum()).add((op * go).sum()).add((w * gw).sum()).backward()
    a2 = (sig * scale).to(DEV).requires_grad_(True)
    b2 = rgb.to(DEV).requires_grad_(True)
    f2, dep2, op2, w2 = ops.composite(a2, b2, z.to(DEV), d.to(DEV), _cfg(spec_o, 0.2), noise.to(DEV), bg.to(DEV) if with_bg else None)
    close(f2, f.detach(), 1e-5, 1e-6, "features")
    (f2 * gf.to(DEV)).sum().add((dep2 * gd.to(DEV)).sum()).add((op2 * go.to(DEV)).sum()).add((w2 * gw.to(DEV)).sum()).backward()
    close(b2.grad, b.grad, 1e-4, 1e-6, "d_rgb")
    close(a2.grad, a.grad, 1e-4, 2e-5 * float(a.grad.abs().max()), "d_sigma")


def test_composite_errors():
    from yanerf.pipelines.renderers.multipass_emission_absorpsion_renderer import EmissionAbsorptionRaymarcher

    m = EmissionAbsorptionRaymarcher(bg_color=(0.0, 0.0))
    x = torch.zeros(2, 3, 5, 1, device=DEV)
    with pytest.raises(ValueError, match="Wrong number of background color channels"):
        m(x, torch.zeros(2, 3, 5, 3, device=DEV), {}, torch.zeros(2, 3, 5, device=DEV), torch.ones(2, 3, 3, device=DEV))
    with pytest.raises(ValueError):
        m(torch.zeros(2, 3, 5, 2, device=DEV), torch.zeros(2, 3, 5, 3, device=DEV), {}, torch.zeros(2, 3, 5, device=DEV),
          torch.ones(2, 3, 3, device=DEV))


# --------------------------------------------------------------------------- ray sampler
def test_ray_bundle_golden(golden):
    from yanerf import ops

    g = golden("raysampler")
    poses, focal = T(g["poses"]).to(DEV), T(g["focal"]).to(DEV)
    B, H, W, P = 2, 6, 10, 16
    depths = O.depth_linspace(0.5, 3.0, P).to(DEV)
    o, d, z, xy = ops.ray_bundle(poses, focal, None, depths, None, H * W, W, H)
    same(xy.reshape(B, H, W, 2), g["ev_xys"], "xys")
    same(o.reshape(B, H, W, 3), g["ev_origins"], "origins")
    close(d.reshape(B, H, W, 3), g["ev_directions"], 1e-6, 1e-7, "directions")
    same(z.reshape(B, H, W, P), g["ev_lengths"], "lengths")
    pix = T(g["pix"]).to(DEV)
    xy_in = torch.stack((pix % W, pix // W), dim=-1).float()
    o, d, z, xy = ops.ray_bundle(poses, focal, xy_in, depths, T(g["u"])[:, :, 0].contiguous().to(DEV), 7, W, H)
    same(xy[:, :, None], g["tr_xys"], "train xys")
    close(d[:, :, None], g["tr_directions"], 1e-6, 1e-7, "train directions")
    same(z[:, :, None], g["tr_lengths"], "stratified lengths")


# --------------------------------------------------------------------------- MLP forward
MLPS = {
    "lego": (O.MLPSpec(), dict()),
    "small": (O.MLPSpec(n_layers=5, input_skips=(2,), n_harmonic_functions_xyz=8, n_hidden_neurons_xyz=64,
                        n_harmonic_functions_dir=4, n_hidden_neurons_dir=32),
              dict(n_layers=5, input_skips=[2], n_harmonic_functions_xyz=8, n_hidden_neurons_xyz=64, n_hidden_neurons_dir=32)),
}


# shallow networks: the data-gradient chain degenerates (with one trunk layer only the folded colour step is left)
MLPS_SHALLOW = {
    "one": (O.MLPSpec(n_layers=1, input_skips=(), n_hidden_neurons_xyz=256),
            dict(n_layers=1, input_skips=[])),
    "two": (O.MLPSpec(n_layers=2, input_skips=(1,), n_hidden_neurons_xyz=128, n_hidden_neurons_dir=64),
            dict(n_layers=2, input_skips=[1], n_hidden_neurons_xyz=128, n_hidden_neurons_dir=64)),
}


def _build_mlp(name, seed, gain, dtype):
    from yanerf.pipelines.models import MODELS
    from tools.testing import LEGO_MLP

    spec, over = {**MLPS, **MLPS_SHALLOW}[name]
    mlp = MODELS.build({**LEGO_MLP, **over})
    mlp.set_operand_dtype(dtype)  # inference and training
    sd = syn.synth_mlp_state(spec.param_shapes(), seed, gain)
    mlp.load_state_dict(sd)
    return mlp.to(DEV), spec, sd


@pytest.mark.parametrize("dtype", ["bf16", "fp16"])
@pytest.mark.parametrize("name", list(MLPS))
@pytest.mark.parametrize("gain,seed", [(1.0, 7), (3.0, 8)])
def test_mlp_forward_golden(golden, name, gain, seed, dtype):
    g = golden("mlp")
    mlp, spec, _ = _build_mlp(name, seed, gain, dtype)
    key = f"{name}_g{int(gain)}"
    o, d, z = (T(g[f"{key}_{k}"]).to(DEV) for k in ("o", "d", "z"))
    with torch.no_grad():
        out = mlp(o, d, z)
    rgb_ref, dens_ref = T(g[key + "_rgb"]), T(g[key + "_density"])
    assert out["rays_features"].shape == rgb_ref.shape and out["rays_densities"].shape == dens_ref.shape
    err_rgb = float((out["rays_features"].cpu() - rgb_ref).abs().max())
    scale = max(1.0, float(dens_ref.abs().max()))
    err_den = float((out["rays_densities"].cpu() - dens_ref).abs().max()) / scale
    print(f"{key} {dtype}: rgb max abs {err_rgb:.2e}; density max abs/scale {err_den:.2e} (scale {scale:.2f})")
    # Stated bound: max abs <= 2e-3 on rgb and raw density for fp16 operands (the default) at lego.yml scale
    # (gain 1); bf16 operands (training) are allowed 5e-3 on the raw density.  gain 3 is a stress case: ten
    # layers of 3x weights push raw densities to O(1e3) and saturate the colour sigmoid, so there the density
    # bound is relative to that scale and rgb is checked through its mean error.
    tol = 2e-3 if dtype == "fp16" else 5e-3
    if gain == 1.0:
        assert err_rgb <= 2e-3, err_rgb
        assert err_den <= tol, err_den
    else:
        assert err_den <= 4 * tol, err_den
        mean_rgb = float((out["rays_features"].cpu() - rgb_ref).abs().mean())
        assert mean_rgb <= 4 * tol, mean_rgb


@pytest.mark.parametrize("R,P", [(3, 64), (257, 192), (1, 1), (130, 7)])
def test_mlp_forward_shapes_and_tails(R, P):
    """Ragged sizes: tiles that straddle rays, a partial last tile, an odd number of tiles."""
    mlp, spec, sd = _build_mlp("lego", 11, 1.0, "fp16")
    rs = np.random.RandomState(R + P)
    o = T((rs.uniform(-0.2, 0.2, size=(R, 3)) + np.array([0, 0, -4.0])).astype(np.float32))
    d = T((rs.uniform(-0.4, 0.4, size=(R, 3)) + np.array([0, 0, 1.0])).astype(np.float32))
    z = T(np.sort(2 + 4 * rs.uniform(size=(R, P)), axis=-1).astype(np.float32))
    with torch.no_grad():
        out = mlp(o.to(DEV)[None], d.to(DEV)[None], z.to(DEV)[None])
        dens_ref, rgb_ref = O.mlp_forward(sd, spec, o, d, z)
    close(out["rays_features"][0], rgb_ref, 0, 2e-3, "rgb")
    close(out["rays_densities"][0, ..., 0], dens_ref, 0, 2e-3 * max(1.0, float(dens_ref.abs().max())), "density")


def test_mlp_unsupported_configs_fail_loudly():
    from yanerf.pipelines.models import MODELS
    from tools.testing import LEGO_MLP

    with pytest.raises(NotImplementedError):
        MODELS.build({**LEGO_MLP, "input_dir": False})
    with pytest.raises(NotImplementedError):
        mlp = MODELS.build({**LEGO_MLP, "n_harmonic_functions_xyz": 12}).to(DEV)
        mlp(torch.zeros(1, 2, 3, device=DEV), torch.ones(1, 2, 3, device=DEV), torch.ones(1, 2, 4, device=DEV))


def test_mlp_latent_codes_golden_and_gradients(golden):
    """`latent_dim > 0` (nerf_mlp.py:158-170, 324-335; the reference's tests/test_models.py and test_pipeline.py:37-64 use
    it): a per-image global code appended to the xyz embedding.  The code is constant over an image, so it enters the
    kernels as per-image effective biases; outputs against the reference (golden, B = 2 images with different codes),
    gradients w.r.t. the code, the code columns of the weights and a bias against autograd through the oracle."""
    from yanerf.pipelines.models import MODELS
    from tools.testing import LEGO_MLP

    g = golden("mlp")
    spec = O.MLPSpec(latent_dim=5)
    sd = syn.synth_mlp_state(spec.param_shapes(), 7, 1.0)
    mlp = MODELS.build({**LEGO_MLP, "latent_dim": 5})
    mlp.load_state_dict(sd)
    mlp = mlp.to(DEV)
    o, d, z, codes = (T(g[f"latent_g1_{k}"]) for k in ("o", "d", "z", "codes"))
    with torch.no_grad():
        out = mlp(o.to(DEV), d.to(DEV), z.to(DEV), global_codes=codes.to(DEV))
    assert out["rays_features"].shape == g["latent_g1_rgb"].shape
    e_rgb = float((out["rays_features"].cpu() - T(g["latent_g1_rgb"])).abs().max())
    e_den = float((out["rays_densities"].cpu() - T(g["latent_g1_density"])).abs().max())
    print(f"latent codes: rgb max abs {e_rgb:.2e}, raw density max abs {e_den:.2e}")
    assert e_rgb <= 2e-3 and e_den <= 2e-3
    with pytest.raises(ValueError):
        mlp(o.to(DEV), d.to(DEV), z.to(DEV))  # a latent network needs its codes (nerf_mlp.py:163-164, 179-183)
    with pytest.raises(ValueError):
        mlp(o.to(DEV), d.to(DEV), z.to(DEV), global_codes=torch.zeros(2, 3, device=DEV))
    # gradients
    rs = np.random.RandomState(4)
    B, n, P = z.shape[0], z.shape[1], z.shape[-1]
    gd = T(rs.standard_normal(size=(B, n, 1, P, 1)).astype(np.float32))
    gc = T(rs.standard_normal(size=(B, n, 1, P, 3)).astype(np.float32))
    ps = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    c_ref = codes.clone().requires_grad_(True)
    rows = c_ref[:, None, :].expand(-1, n, -1).reshape(-1, 5)
    dens_ref, rgb_ref = O.mlp_forward(ps, spec, o.reshape(-1, 3), d.reshape(-1, 3), z.reshape(-1, P), rows)
    ((dens_ref * gd.reshape(-1, P)).sum() + (rgb_ref * gc.reshape(-1, P, 3)).sum()).backward()
    c_dev = codes.to(DEV).requires_grad_(True)
    out = mlp(o.to(DEV), d.to(DEV), z.to(DEV), global_codes=c_dev)
    ((out["rays_densities"] * gd.to(DEV)).sum() + (out["rays_features"] * gc.to(DEV)).sum()).backward()

    def cos(a, b):
        a, b = a.detach().cpu().double().reshape(-1), b.detach().double().reshape(-1)
        return float((a * b).sum() / (a.norm() * b.norm()).clamp_min(1e-30)), float(a.norm() / b.norm().clamp_min(1e-30))

    got = dict(mlp.named_parameters())
    for name, a, b in (("codes", c_dev.grad, c_ref.grad),
                       ("layer 0 code columns", got["xyz_encoder.mlp.0.0.weight"].grad[:, 63:], ps["xyz_encoder.mlp.0.0.weight"].grad[:, 63:]),
                       ("skip layer code columns", got["xyz_encoder.mlp.5.0.weight"].grad[:, 319:], ps["xyz_encoder.mlp.5.0.weight"].grad[:, 319:]),
                       ("layer 0 bias", got["xyz_encoder.mlp.0.0.bias"].grad, ps["xyz_encoder.mlp.0.0.bias"].grad),
                       ("layer 3 weight", got["xyz_encoder.mlp.3.0.weight"].grad, ps["xyz_encoder.mlp.3.0.weight"].grad)):
        c, ratio = cos(a, b)
        print(f"latent codes: d {name}: cos {c:.5f}, norm ratio {ratio:.4f}")
        assert c >= 0.99 and abs(ratio - 1) <= 0.05, (name, c, ratio)


# --------------------------------------------------------------------------- MLP backward
def _ste_round(x, dt):
    """Round to the tensor-core operand type, identity gradient."""
    return x + (x.to(dt).float() - x).detach()


def mlp_forward_operand_rounded(params, spec, origins, directions, lengths, dt):
    """The oracle's `mlp_forward` (nerf_mlp.py:117-177) with the kernel's documented operand rounding inserted:
    embedding, trunk / intermediate weights+biases and every layer output feeding a tensor-core layer are rounded
    to `dt` -- including the colour hidden activations and W2 of the colour head, which is a tensor-core layer too;
    the density head, the per-ray direction bias and the colour head's bias stay fp32.  The intermediate layer is
    linear and merged into the colour hidden layer: ONE rounded weight matrix W_c[:, :H] W_i and bias W_c[:, :H] b_i, the
    intermediate output itself is never rounded."""
    import torch.nn.functional as F

    r = lambda t: _ste_round(t, dt)
    pts = origins[:, None, :] + lengths[:, :, None] * directions[:, None, :]
    emb = r(O.harmonic_embedding(pts, spec.n_harmonic_functions_xyz))
    y = emb
    for li in range(spec.n_layers):
        if li in spec.input_skips:
            y = torch.cat((y, emb), dim=-1)
        pre = torch.relu(F.linear(y, r(params[f"xyz_encoder.mlp.{li}.0.weight"]), r(params[f"xyz_encoder.mlp.{li}.0.bias"])))
        y = r(pre)
    raw_density = F.linear(pre, params["density_layer.weight"], params["density_layer.bias"])[..., 0]
    demb = O.harmonic_embedding(F.normalize(directions, dim=-1), spec.n_harmonic_functions_dir)
    h = spec.n_hidden_neurons_xyz
    wc = params["color_layer.0.weight"]
    w_ci = r(wc[:, :h] @ params["intermediate_linear.weight"])
    b_ci = r(wc[:, :h] @ params["intermediate_linear.bias"])
    hid = r(torch.relu(F.linear(y, w_ci, b_ci) + (F.linear(demb, wc[:, h:], params["color_layer.0.bias"]))[:, None, :]))
    rgb = torch.sigmoid(F.linear(hid, r(params["color_layer.2.weight"]), params["color_layer.2.bias"]))
    return raw_density, rgb


@pytest.mark.parametrize("name,R,P", [("lego", 70, 64), ("lego", 33, 192), ("small", 19, 24), ("lego", 2, 3),
                                      ("one", 700, 64), ("two", 300, 64)])  # "one": 175 tile pairs, some CTAs walk two
@pytest.mark.parametrize("dtype", ["bf16", "fp16"])
def test_mlp_backward_vs_oracle_autograd(name, R, P, dtype):
    """Parameter gradients of the tcgen05 data-/weight-gradient kernels vs torch autograd.

    (a) against the oracle with the kernel's operand rounding inserted (straight-through): this isolates the
        backward implementation; per-tensor relative L2 error <= 2e-2, cosine >= 0.9995;
    (b) against the plain fp32 oracle: the difference is dominated by ReLU units whose pre-activation sign flips
        under 16-bit operands (relative L2 ~ sqrt(fraction flipped)), so only cosine >= 0.99 / rel <= 0.2 holds.
    """
    mlp, spec, sd = _build_mlp(name, 13, 1.0, dtype)
    dt = torch.bfloat16 if dtype == "bf16" else torch.float16
    rs = np.random.RandomState(R * 7 + P)
    o = T((rs.uniform(-0.2, 0.2, size=(R, 3)) + np.array([0, 0, -4.0])).astype(np.float32))
    d = T((rs.uniform(-0.4, 0.4, size=(R, 3)) + np.array([0, 0, 1.0])).astype(np.float32))
    z = T(np.sort(2 + 4 * rs.uniform(size=(R, P)), axis=-1).astype(np.float32))
    gd = T(rs.standard_normal(size=(R, P)).astype(np.float32))
    gc = T(rs.standard_normal(size=(R, P, 3)).astype(np.float32))
    refs = {}
    for tag in ("rounded", "fp32"):
        ps = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        if tag == "fp32":
            dens_ref, rgb_ref = O.mlp_forward(ps, spec, o, d, z)
        else:
            dens_ref, rgb_ref = mlp_forward_operand_rounded(ps, spec, o, d, z, dt)
        ((dens_ref * gd).sum() + (rgb_ref * gc).sum()).backward()
        refs[tag] = ps
    out = mlp(o.to(DEV)[None], d.to(DEV)[None], z.to(DEV)[None])
    # forward of the training-mode kernel (stash on) against the fp32 oracle
    assert float((out["rays_features"][0].detach().cpu() - rgb_ref.detach()).abs().max()) <= (5e-3 if dtype == "bf16" else 2e-3)
    assert float((out["rays_densities"][0, ..., 0].detach().cpu() - dens_ref.detach()).abs().max()) <= (1e-2 if dtype == "bf16" else 2e-3)
    ((out["rays_densities"][0, ..., 0] * gd.to(DEV)).sum() + (out["rays_features"][0] * gc.to(DEV)).sum()).backward()
    worst = {"rounded": 0.0, "fp32": 0.0}
    for k, p in mlp.named_parameters():
        g = p.grad.detach().cpu().double().reshape(-1)
        for tag, (tol, mincos) in (("rounded", (2e-2, 0.9995)), ("fp32", (0.2, 0.99))):
            gr = refs[tag][k].grad.double().reshape(-1)
            rel = float((g - gr).norm() / gr.norm().clamp_min(1e-12))
            cos = float(torch.dot(g, gr) / (g.norm() * gr.norm()).clamp_min(1e-30))
            worst[tag] = max(worst[tag], rel)
            assert rel <= tol and cos >= mincos, f"{k} vs {tag}: rel L2 {rel:.3e}, cos {cos:.6f}, |ref| {float(gr.norm()):.3e}"
    print(f"{name} R={R} P={P} {dtype}: worst rel L2 gradient error {worst['rounded']:.2e} (operand-rounded oracle), "
          f"{worst['fp32']:.2e} (fp32 oracle)")


def test_composite_generic_and_blocked_paths_agree():
    """P in {64,128,192} with 8-byte aligned rows takes the lane-blocked kernels, anything else the chunked ones:
    feed the same rays through both (the second copy is deliberately misaligned by one float)."""
    from yanerf import ops

    rs = np.random.RandomState(3)
    R, P = 300, 128
    mk = lambda *s: torch.from_numpy(rs.uniform(size=s).astype(np.float32)).to(DEV)
    sig, rgb, z, d = mk(R, P) * 4 - 1, mk(R, P, 3), torch.sort(2 + 4 * mk(R, P), dim=-1)[0], mk(R, 3) - 0.5
    cfg = ops.march_cfg(1e10, 1e-6, 0.0, False, False, (0.1, 0.2, 0.3))

    def misaligned(t):
        buf = torch.empty(t.numel() + 1, device=DEV)
        buf[1:].copy_(t.reshape(-1))
        return buf[1:].view(t.shape)

    outs = []
    for f in (lambda t: t, misaligned):
        a, b = f(sig).detach().requires_grad_(True), f(rgb).detach().requires_grad_(True)
        feats, dep, op, w = ops.composite(a, b, f(z), d, cfg)
        (feats.sum() + 0.3 * dep.sum() + (w * w).sum()).backward()
        outs.append((feats, dep, op, w, a.grad, b.grad))
    for x, y, name in zip(outs[0], outs[1], ("features", "depths", "opacities", "weights", "d_sigma", "d_rgb")):
        close(x, y.detach().cpu(), 2e-5, 2e-6, name)


@pytest.mark.parametrize("P,n", [(64, 128), (64, 64), (128, 128), (128, 64), (192, 128), (192, 64)])
def test_refiner_fast_and_generic_kernels_agree_bitwise(P, n):
    """The lane-blocked fast path (aligned rows, lego / fern shapes) and the generic kernel (forced here by a
    one-float misalignment of the output) must produce identical bits, random and deterministic draws, sorted and
    unsorted input depths."""
    from yanerf import ops

    rs = np.random.RandomState(P + n)
    R = 3000
    z = np.sort(2 + 4 * rs.uniform(size=(R, P)).astype(np.float32), axis=-1)
    z[5] = z[5, ::-1].copy()  # one descending row: exercises the unsorted-input fallback
    w = rs.uniform(size=(R, P)).astype(np.float32) ** 8
    w[rs.uniform(size=w.shape) < 0.5] = 0.0
    u = np.minimum(rs.uniform(size=(R, n)).astype(np.float32), np.float32(1 - 2 ** -24))
    zt, wt, ut = T(z).to(DEV), T(w).to(DEV), T(u).to(DEV)
    for uu in (None, ut):
        fast, inds_f, _ = ops.sample_pdf_merge(zt, wt, n, uu, want_inds=True)
        buf = torch.empty(R * (P + n) + 1, device=DEV)
        out = buf[1:].view(R, P + n)
        inds_g = torch.empty(R, n, dtype=torch.int64, device=DEV)
        flag = torch.zeros(1, dtype=torch.int32, device=DEV)
        uarg, stride = (ops.det_draws(n, zt.device), 0) if uu is None else (uu, n)
        ops._call("yn_sample_pdf_merge", ops.N.ptr(zt), ops.N.ptr(wt), ops.N.ptr(uarg), stride, None, 0,
                  ctypes_ptr(out), ops.N.ptr(inds_g, torch.int64), ops.N.ptr(flag, torch.int32), R, P, n, 1, ops.STREAM,
                  device=zt.device)
        same(inds_f, inds_g.cpu(), "inds")
        same(fast, out.cpu(), "lengths")
        ref, ref_inds = O.refine_lengths(T(z), T(w), n, None if uu is None else T(u))
        same(inds_f, ref_inds, "inds vs oracle")
        same(fast, ref, "lengths vs oracle")


def ctypes_ptr(t):
    import ctypes

    return ctypes.c_void_p(t.data_ptr())


def test_pixel_sampler_distinct_uniform():
    """yn_sample_pixels: n distinct in-range pixels per image, different per image and per seed, uniform marginals."""
    from yanerf import ops

    H, W, B, n = 37, 53, 3, 700
    seed = torch.tensor([12345], dtype=torch.int64, device=DEV)
    idx, xy = ops.sample_pixels(seed, B, n, W, H)
    assert idx.shape == (B, n) and int(idx.min()) >= 0 and int(idx.max()) < H * W
    for b in range(B):
        assert idx[b].unique().numel() == n
    assert not torch.equal(idx[0], idx[1])
    assert torch.equal(xy[..., 0].long() + W * xy[..., 1].long(), idx)
    idx2, _ = ops.sample_pixels(seed + 1, B, n, W, H)
    assert not torch.equal(idx, idx2)
    full, _ = ops.sample_pixels(seed, 1, H * W, W, H)  # n == H*W: a permutation
    assert torch.equal(full[0].sort()[0], torch.arange(H * W, device=DEV))
    # uniformity: 2000 draws of 64 pixels from a 16x16 image, chi-square over the 256 cells
    counts = torch.zeros(256, device=DEV)
    for s in range(200):
        d, _ = ops.sample_pixels(seed + 100 + s, 10, 64, 16, 16)
        counts += torch.bincount(d.reshape(-1), minlength=256).float()
    expected = 200 * 10 * 64 / 256
    chi2 = float(((counts - expected) ** 2 / expected).sum())
    assert chi2 < 255 + 6 * (2 * 255) ** 0.5, chi2  # mean 255, sigma ~22.6


# --------------------------------------------------------------------------- fused Adam (row A)
@pytest.mark.parametrize("world", [1, 8])
def test_adam_kernel_vs_torch_and_oracle(world):
    """`yn_adam_step` (host step / lr) and `yn_adam_step_dev` (step / lr in device memory) against torch.optim.Adam
    (the optimizer scripts/run.py:159 builds) AND the oracle's `adam_step`, 10 steps on random flat buffers with a
    changing learning rate; `grad_scale = 1 / world` models the mean of the DDP all-reduce (the kernels see the SUM of
    the ranks' gradients).  Bound: rtol 1e-6 plus a few fp32 ulps of the largest gradient that entered the element (the
    first moment is a signed sum: torch's `lerp_`, the oracle's `mul_().add_()` and the kernel's fused form round
    differently by one ulp of the operands, which is unbounded RELATIVE to an element that cancels to near zero)."""
    from yanerf import ops

    n = 100_003
    rs = np.random.RandomState(11 + world)
    p0 = T((0.01 * rs.standard_normal(n)).astype(np.float32))  # small, so that one ulp of p is far below one update
    grads = [T((rs.standard_normal(n) * 10.0 ** rs.uniform(-6, 0, size=n)).astype(np.float32)) for _ in range(10)]
    lrs = [5e-4 * 0.9 ** k for k in range(10)]
    # torch.optim.Adam on the CPU
    pt = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([pt], lr=lrs[0])
    # the oracle's restatement
    po, mo, vo = p0.clone(), torch.zeros(n), torch.zeros(n)
    # the two kernels
    pk, mk, vk = p0.clone().to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    pd, md, vd = p0.clone().to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    state = torch.zeros(2, device=DEV)
    for k, (g, lr) in enumerate(zip(grads, lrs)):
        opt.param_groups[0]["lr"] = lr
        pt.grad = g.clone()
        opt.step()
        O.adam_step(po, g, mo, vo, k + 1, lr)
        gsum = (g * world).to(DEV)  # what the all-reduce (SUM) leaves in the flat gradient buffer
        ops.adam_step(pk, gsum, mk, vk, lr, k + 1, grad_scale=1.0 / world)
        state[0], state[1] = float(k + 1), lr
        ops.adam_step_dev(pd, gsum, md, vd, state, grad_scale=1.0 / world)
    st = opt.state[pt]
    gmax = torch.stack(grads).abs().max(dim=0)[0].double()
    ulp = 2.0 ** -23

    def close_ulp(got, ref, scale, what):
        got, ref = got.detach().cpu().double(), ref.detach().double()
        bad = (got - ref).abs() > 1e-6 * ref.abs() + 4 * ulp * scale
        assert not bad.any(), f"{what}: {int(bad.sum())}/{bad.numel()} off, max abs {float((got - ref).abs().max()):.3e}"

    for name, (p_, m_, v_) in (("host-step kernel", (pk, mk, vk)), ("device-step kernel", (pd, md, vd))):
        for ref_name, (pr, mr, vr) in (("torch.optim.Adam", (pt.detach(), st["exp_avg"], st["exp_avg_sq"])), ("oracle", (po, mo, vo))):
            close_ulp(m_, mr, gmax, f"{name} exp_avg vs {ref_name}")
            close_ulp(v_, vr, gmax * gmax, f"{name} exp_avg_sq vs {ref_name}")
            # a parameter is the sum of 10 updates of size ~lr: compare the travelled distance, not the (O(1)) value;
            # one ulp of the parameter itself (|p| ~ 1) is the floor of every single update
            close(p_ - p0.to(DEV), pr - p0, 1e-5, 10 * ulp * (float(p0.abs().max()) + 5e-3), f"{name} parameter update vs {ref_name}")
    assert torch.equal(pk, pd) or float((pk - pd).abs().max()) <= 1e-9


# --------------------------------------------------------------------------- in-kernel draws
def test_device_rng_matches_explicit_draws_bitwise():
    """Every randomised kernel has two draw sources: an explicit tensor (the parity tests replay the reference's draws
    through it) and the in-kernel Philox stream.  `yn_rng_fill` writes out the stream; feeding it back through the
    explicit pointers must give the same bits as the in-kernel path: stratified depths, density noise forward AND backward
    (blocked and generic kernels), inverse-CDF uniforms (fast and generic kernels)."""
    from yanerf import ops

    rng = ops.DeviceRng(DEV, seed=1234)
    ops.step_begin(rng, None)
    ops.step_begin(rng, None)  # current step = 1
    rs = np.random.RandomState(3)
    # --- stratified jitter + pixel pick
    B, n, P, H, W = 2, 300, 64, 40, 50
    poses, focal = syn.synth_camera(B, seed=1).to(DEV), torch.full((B,), 33.0, device=DEV)
    depths = O.depth_linspace(2.0, 6.0, P).to(DEV)
    idx, xys, o, d, z = ops.train_rays(rng, poses, focal, depths, True, n, W, H)
    for b in range(B):
        assert idx[b].unique().numel() == n and int(idx[b].min()) >= 0 and int(idx[b].max()) < H * W
    assert torch.equal(xys[..., 0].long() + W * xys[..., 1].long(), idx)
    u = ops.rng_fill(rng, ops.DeviceRng.SITE_STRATIFIED, B * n, P, normal=False).reshape(B, n, P)
    o2, d2, z2, _ = ops.ray_bundle(poses, focal, xys, depths, u, n, W, H)
    same(z, z2.cpu(), "stratified depths"); same(o, o2.cpu(), "origins"); same(d, d2.cpu(), "directions")
    assert 0.0 <= float(u.min()) and float(u.max()) < 1.0 and abs(float(u.mean()) - 0.5) < 0.01
    assert abs(float(u.var()) - 1 / 12) < 0.005
    # --- density noise: in-kernel vs explicit, forward and backward, P = 64 / 192 (blocked) and 37 (generic)
    for Pn in (64, 192, 37):
        R = 500
        sig = T((rs.standard_normal(size=(R, Pn)) * 0.5).astype(np.float32)).to(DEV)
        rgb = T(rs.uniform(size=(R, Pn, 3)).astype(np.float32)).to(DEV)
        zz = T(np.sort(2 + 4 * rs.uniform(size=(R, Pn)), axis=-1).astype(np.float32)).to(DEV)
        dd = T(rs.standard_normal(size=(R, 3)).astype(np.float32)).to(DEV)
        cfg = ops.march_cfg(1e10, 1e-6, 0.3, False, False, (0.0, 0.0, 0.0))
        site = ops.DeviceRng.SITE_NOISE + 1
        noise = ops.rng_fill(rng, site, R, Pn, normal=True)
        outs = []
        for kw in (dict(noise=noise), dict(rng=rng, site=site)):
            a, b_ = sig.clone().requires_grad_(True), rgb.clone().requires_grad_(True)
            f, dep, op, w = ops.composite(a, b_, zz, dd, cfg, **kw)
            (f.sum() + (w * w).sum()).backward()
            outs.append((f, dep, w, a.grad, b_.grad))
        for x, y, name in zip(outs[0], outs[1], ("features", "depths", "weights", "d_sigma", "d_rgb")):
            same(y, x.detach().cpu(), f"P={Pn} {name}")
        assert abs(float(noise.mean())) < 0.02 and abs(float(noise.var()) - 1.0) < 0.03
        assert abs(float((noise ** 4).mean()) - 3.0) < 0.3  # kurtosis of a normal
    # --- inverse-CDF uniforms: fast kernel (64 -> +128) and generic kernel (50 -> +20)
    for Pn, nn in ((64, 128), (50, 20)):
        R = 700
        zz = T(np.sort(2 + 4 * rs.uniform(size=(R, Pn)), axis=-1).astype(np.float32)).to(DEV)
        ww = T((rs.uniform(size=(R, Pn)) ** 4).astype(np.float32)).to(DEV)
        site = ops.DeviceRng.SITE_PDF
        uu = ops.rng_fill(rng, site, R, nn, normal=False)
        a, ia, _ = ops.sample_pdf_merge(zz, ww, nn, uu, want_inds=True)
        b_, ib, _ = ops.sample_pdf_merge(zz, ww, nn, None, want_inds=True, rng=rng, site=site)
        same(b_, a.cpu(), f"refined depths P={Pn}"); same(ib, ia.cpu(), f"inds P={Pn}")
        ref, ref_inds = O.refine_lengths(zz.cpu(), ww.cpu(), nn, uu.cpu())
        same(b_, ref, "in-kernel draws vs oracle on the same uniforms"); same(ib, ref_inds, "inds vs oracle")
    # --- a new step gives new draws; the same step gives the same draws
    u_again = ops.rng_fill(rng, ops.DeviceRng.SITE_STRATIFIED, B * n, P, normal=False).reshape(B, n, P)
    assert torch.equal(u, u_again)
    ops.step_begin(rng, None)
    u_next = ops.rng_fill(rng, ops.DeviceRng.SITE_STRATIFIED, B * n, P, normal=False).reshape(B, n, P)
    assert not torch.equal(u, u_next) and abs(float((u * u_next).mean()) - 0.25) < 0.01  # uncorrelated


@pytest.mark.parametrize("H,W,n", [(30, 44, 1000), (300, 400, 100000)])  # 16-block and 128-block (full-image) grids
def test_rgb_loss_and_scatter_kernels_vs_torch(H, W, n):
    """`yn_rgb_loss_fwd/bwd` (GT gather + per-image mse / huber) and `yn_scatter_rays` against the torch ops they replace
    (the reference's sample_grid / _rgb_metrics / scatter_rays_to_image, pipelines/utils.py)."""
    from yanerf import ops
    from yanerf.pipelines.utils import _rgb_metrics, sample_grid, scatter_rays_to_image

    B = 3
    g = torch.Generator().manual_seed(0)
    image = torch.rand(B, H, W, 3, generator=g).to(DEV)
    idx = torch.stack([torch.randperm(H * W, generator=g)[:n] for _ in range(B)]).to(DEV)
    xy = torch.stack((idx % W, idx // W), dim=-1).float()
    pred = torch.rand(B, n, 3, generator=g).to(DEV)
    p1 = pred.clone().requires_grad_(True)
    ref = _rgb_metrics(sample_grid(image, xy[:, :, None], validate=False), p1[:, :, None])
    (ref["rgb_mse"] * torch.tensor([1.0, 2.0, 3.0], device=DEV)).sum().add(ref["rgb_huber"].sum() * 0.5).backward()
    p2 = pred.clone().requires_grad_(True)
    mse, hub = ops.rgb_loss(p2, image, xy)
    (mse * torch.tensor([1.0, 2.0, 3.0], device=DEV)).sum().add(hub.sum() * 0.5).backward()
    close(mse, ref["rgb_mse"].detach().cpu(), 1e-6 if n <= 65536 else 3e-6, 0, "mse")  # (torch's fp32 tree sum vs fp64 here)
    close(hub, ref["rgb_huber"].detach().cpu(), 1e-5, 1e-9, "huber")
    mse_again, _ = ops.rgb_loss(pred, image, xy)
    same(mse_again, mse.detach().cpu(), "rgb_loss is deterministic (fixed-order fold, self-resetting counter)")
    close(p2.grad, p1.grad.cpu(), 1e-5, 1e-10, "d_pred")
    dep, alp = torch.rand(B, n, 1, device=DEV), torch.rand(B, n, 1, device=DEV)
    outs = ops.scatter_rays([pred, dep, alp], xy, H, W)
    for got, src in zip(outs, (pred, dep, alp)):
        assert got.is_contiguous()
        same(got, scatter_rays_to_image(src[:, :, None], xy[:, :, None], H, W).cpu(), "scatter")
