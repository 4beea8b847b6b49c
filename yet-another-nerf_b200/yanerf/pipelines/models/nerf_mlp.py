"""NeRFMLP: harmonic embedding + 8x256 skip trunk + density and view-dependent colour heads.

Same registry name, constructor keywords (including the reference's spellings), forward signature,
output dict and state-dict keys as `yanerf/pipelines/models/nerf_mlp.py:12-183`, so reference configs and
checkpoints load unchanged.  The torch sub-modules below only HOLD the parameters under the reference's
names; the arithmetic of `forward` is the fused sm_100a kernel chain (`yn_mlp_dirbias`, `yn_mlp_fwd`,
`yn_mlp_bwd`): there is no torch implementation of the network in this package.
"""
from __future__ import annotations

import math
import os
from typing import List, Optional

import torch

from ... import _native as N
from ... import ops
from ...utils.logging import get_logger

from .builder import MODELS


class LinearWithRepeat(torch.nn.Module):
    """Parameter holder of the colour hidden layer (models/utils.py:135-211): weight [out, n1 + n2]."""

    def __init__(self, in_features: int, out_features: int) -> None:
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = torch.nn.Parameter(torch.empty(out_features, in_features))
        self.bias = torch.nn.Parameter(torch.empty(out_features))
        torch.nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        bound = 1 / math.sqrt(in_features)
        torch.nn.init.uniform_(self.bias, -bound, bound)


class MLPWithInputSkips(torch.nn.Module):
    """Parameter holder of the trunk (nerf_mlp.py:186-289): `mlp.{l}.0` = Linear, `mlp.{l}.1` = ReLU.
    Inner width is always `hidden_dim` = 256; only the last layer emits `output_dim` (nerf_mlp.py:88-95)."""

    def __init__(self, n_layers: int, input_dim: int, output_dim: int, skip_dim: int, input_skips, hidden_dim: int = 256):
        super().__init__()
        layers = []
        for li in range(n_layers):
            din = input_dim if li == 0 else (hidden_dim + skip_dim if li in input_skips else hidden_dim)
            dout = hidden_dim if li + 1 < n_layers else output_dim
            lin = torch.nn.Linear(din, dout)
            torch.nn.init.xavier_uniform_(lin.weight.data)
            layers.append(torch.nn.Sequential(lin, torch.nn.ReLU(True)))
        self.mlp = torch.nn.ModuleList(layers)
        self._input_skips = set(input_skips)


_FMT = {"fp16": N.FMT_FP16, "bf16": N.FMT_BF16}


def _default_fmt(training: bool) -> int:
    """Tensor-core operand type: fp16 (10-bit mantissa, 4x tighter than bf16; NeRF activations are far inside its range)
    for inference AND training, so that the training forward is the evaluation forward.  The 16-bit gradients of a
    backward call are carried times one power of two chosen from the largest incoming gradient (`grad_scale`,
    csrc/mlp_bwd.cu) -- exact, and it keeps them inside fp16's exponent range.  `YANERF_MLP_TRAIN_DTYPE=bf16` (or
    `set_operand_dtype`) selects bf16 activations and gradients instead (fp32's exponent range, no scaling)."""
    if training:
        return _FMT[os.environ.get("YANERF_MLP_TRAIN_DTYPE", "fp16").lower()]
    return _FMT[os.environ.get("YANERF_MLP_DTYPE", "fp16").lower()]


@MODELS.register_module()
class NeRFMLP(torch.nn.Module):
    def __init__(
        self,
        n_layers: int = 8,
        input_skips: List[int] = [5],
        n_harmonic_functions_xyz: int = 10,
        harmonic_functions_xyz_append_intput: bool = True,
        n_hidden_neurons_xyz: int = 256,
        n_harmonic_functions_dir: int = 4,
        harmonic_functions_dir_append_intput: bool = True,
        n_hidden_neurons_dir: int = 128,
        latent_dim: int = 0,
        input_xyz: bool = True,
        input_dir: bool = True,
        color_dim: int = 3,
        nerf_paper_v1=False,
    ) -> None:
        super().__init__()
        self.logger = get_logger(__name__)
        if not input_xyz and latent_dim <= 0:
            raise ValueError("The latent dimension has to be > 0 if xyz is not input!")
        unsupported = []
        if not input_xyz or not input_dir:
            unsupported.append("input_xyz / input_dir = False")
        if not harmonic_functions_xyz_append_intput or not harmonic_functions_dir_append_intput:
            unsupported.append("harmonic embedding without the appended input")
        if nerf_paper_v1:
            unsupported.append("nerf_paper_v1 extra colour layers")
        if unsupported:
            raise NotImplementedError(
                "the sm_100a NeRF-MLP kernel covers the lego.yml / fern.yml architecture family only; unsupported: "
                + ", ".join(unsupported)
            )
        self.n_layers = n_layers
        self.input_skips = list(input_skips)
        self.n_harmonic_functions_xyz = n_harmonic_functions_xyz
        self.harmonic_functions_xyz_append_intput = harmonic_functions_xyz_append_intput
        self.n_hidden_neurons_xyz = n_hidden_neurons_xyz
        self.n_harmonic_functions_dir = n_harmonic_functions_dir
        self.harmonic_functions_dir_append_intput = harmonic_functions_dir_append_intput
        self.n_hidden_neurons_dir = n_hidden_neurons_dir
        self.latent_dim = latent_dim
        self.input_xyz, self.input_dir = input_xyz, input_dir
        self.color_dim = color_dim

        embed_xyz = 3 * (2 * n_harmonic_functions_xyz + 1)
        embed_dir = 3 * (2 * n_harmonic_functions_dir + 1)
        self._embed_xyz = embed_xyz
        if latent_dim > 0:
            self.logger.info(f"Model, use `global_codes`, latent_dim = {latent_dim}.")  # nerf_mlp.py:47-48
        # the global code is appended to the xyz embedding (nerf_mlp.py:85-86, 324-335): layer 0 and every skip layer
        # see embed_xyz + latent_dim input columns
        self.xyz_encoder = MLPWithInputSkips(n_layers, embed_xyz + latent_dim, n_hidden_neurons_xyz, embed_xyz + latent_dim,
                                             self.input_skips)
        self.intermediate_linear = torch.nn.Linear(n_hidden_neurons_xyz, n_hidden_neurons_xyz)
        torch.nn.init.xavier_uniform_(self.intermediate_linear.weight.data)
        self.density_layer = torch.nn.Linear(n_hidden_neurons_xyz, 1)
        torch.nn.init.xavier_uniform_(self.density_layer.weight.data)
        self.density_layer.bias.data[:] = 0.0
        self.color_layer = torch.nn.Sequential(
            LinearWithRepeat(n_hidden_neurons_xyz + embed_dir, n_hidden_neurons_dir),
            torch.nn.ReLU(True),
            torch.nn.Linear(n_hidden_neurons_dir, color_dim),
            torch.nn.Sigmoid(),
        )

        skip_mask = 0
        for li in self.input_skips:
            if 0 < li < n_layers:
                skip_mask |= 1 << li
        self._arch_fields = (n_layers, skip_mask, n_harmonic_functions_xyz, n_harmonic_functions_dir,
                             n_hidden_neurons_xyz, n_hidden_neurons_dir, color_dim)
        self._fmt = {False: _default_fmt(False), True: _default_fmt(True)}  # keyed by "needs grad"
        self._plans = {}  # fmt -> [MlpPlan, packed key]
        self._flat_leaf: Optional[torch.Tensor] = None  # set by FusedTrainer: parameters live in one flat buffer
        self._flat_grad: Optional[torch.Tensor] = None
        self._ordered: Optional[List[torch.nn.Parameter]] = None  # cache of ordered_parameters()
        self._flat_nograd = [None, None]  # [parameter key, flat copy] reused by no-grad forwards until a parameter changes

    # ------------------------------------------------------------------ parameter plumbing
    def ordered_parameters(self) -> List[torch.nn.Parameter]:
        """State-dict order = the flat layout `yn_mlp_pack_weights` expects."""
        if self._ordered is None:
            ps: List[torch.nn.Parameter] = []
            for seq in self.xyz_encoder.mlp:
                ps += [seq[0].weight, seq[0].bias]
            ps += [self.intermediate_linear.weight, self.intermediate_linear.bias]
            ps += [self.density_layer.weight, self.density_layer.bias]
            ps += [self.color_layer[0].weight, self.color_layer[0].bias, self.color_layer[2].weight, self.color_layer[2].bias]
            self._ordered = ps
        return self._ordered

    def _apply(self, fn, *args, **kwargs):  # .to() / .cuda() / .float() may swap the Parameter objects
        self._ordered = None
        self._flat_nograd = [None, None]
        return super()._apply(fn, *args, **kwargs)

    def _param_key(self, device) -> tuple:
        return (str(device),) + tuple((p.data_ptr(), p._version) for p in self.ordered_parameters())

    def set_operand_dtype(self, name: str, training: Optional[bool] = None) -> None:
        """'bf16' or 'fp16' tensor-core operands (fp32 accumulation either way) for inference, training or
        (training=None) both."""
        for mode in ((False, True) if training is None else (training,)):
            self._fmt[mode] = _FMT[name]

    def _flat(self) -> torch.Tensor:
        return torch.cat([p.reshape(-1) for p in self.ordered_parameters()])

    def use_flat_parameters(self, flat_leaf: Optional[torch.Tensor], flat_grad: Optional[torch.Tensor]) -> None:
        """The module's parameters are (already) consecutive views of `flat_leaf`'s storage: use it as the autograd
        leaf and let the backward kernels accumulate straight into `flat_grad` (no per-tensor cat / split / add)."""
        if flat_leaf is not None and self.latent_dim > 0:
            raise NotImplementedError("flat-buffer training (FusedTrainer) of a NeRFMLP with latent_dim > 0: the kernels see "
                                      "per-image effective parameters (see `_effective_flat`); train it through autograd")
        if flat_leaf is not None:
            ps = self.ordered_parameters()
            ptr = flat_leaf.data_ptr()
            for p in ps:
                if p.data_ptr() != ptr:
                    raise ValueError("parameters are not consecutive views of the flat buffer")
                ptr += p.numel() * 4
            assert flat_grad is not None and flat_grad.numel() == flat_leaf.numel()
        self._flat_leaf, self._flat_grad = flat_leaf, flat_grad
        self.invalidate_packed_weights()

    def plan_for(self, flat: torch.Tensor, needs_grad: bool, key: Optional[tuple] = None) -> ops.MlpPlan:
        """(Re)pack the tensor-core weight image when the parameters changed since the last call."""
        fmt = self._fmt[bool(needs_grad)]
        if key is None:
            key = self._param_key(flat.device)
        entry = self._plans.get(fmt)
        if entry is None or entry[0].wpack.device != flat.device:
            entry = [ops.MlpPlan.create(N.MlpArch(*self._arch_fields, fmt), flat.device), None]
            self._plans[fmt] = entry
        if key != entry[1]:
            entry[0].pack(flat.detach())
            entry[1] = key
        return entry[0]

    def _effective_flat(self, code: torch.Tensor) -> torch.Tensor:
        """Flat parameter vector of the code-free architecture the kernels implement, for ONE image's global code.
        The code is constant over the image's points (`broadcast_global_code`, nerf_mlp.py:324-335), so in every layer
        that reads the embedding its columns act as a bias: `W [emb | code] + b = W[:, :E] emb + (b + W[:, E:] code)`.
        Built with differentiable torch ops on [dout x latent_dim] slices: autograd carries the kernels' gradient of the
        effective parameters back to the code columns, the biases and the code itself."""
        E, L, ps = self._embed_xyz, self.latent_dim, []
        for li, seq in enumerate(self.xyz_encoder.mlp):
            w, b = seq[0].weight, seq[0].bias
            if li == 0 or li in self.xyz_encoder._input_skips:
                keep = w.shape[1] - L  # [hidden | emb] then the code columns
                assert keep == (E if li == 0 else 256 + E)
                b = b + w[:, keep:] @ code
                w = w[:, :keep]
            ps += [w, b]
        ps += [self.intermediate_linear.weight, self.intermediate_linear.bias, self.density_layer.weight, self.density_layer.bias,
               self.color_layer[0].weight, self.color_layer[0].bias, self.color_layer[2].weight, self.color_layer[2].bias]
        return torch.cat([p.reshape(-1) for p in ps])

    def invalidate_packed_weights(self) -> None:
        """Call after updating parameters behind torch's back (e.g. a fused optimizer kernel)."""
        for entry in self._plans.values():
            entry[1] = None
        self._flat_nograd = [None, None]

    # ------------------------------------------------------------------ forward
    def forward(self, origins: torch.Tensor, directions: torch.Tensor, lengths: torch.Tensor,
                global_codes: Optional[torch.Tensor] = None, **kwargs) -> dict:
        """origins/directions `[B,*sp,3]`, lengths `[B,*sp,P]` -> rays_densities `[B,*sp,P,1]` (raw),
        rays_features `[B,*sp,P,color_dim]`, aux {}."""
        if global_codes is not None:
            global_codes = global_codes.reshape(global_codes.shape[0], -1)  # nerf_mlp.py:160-161
        if (global_codes is None) != (self.latent_dim == 0) or (global_codes is not None and global_codes.shape[-1] != self.latent_dim):
            raise ValueError("The shape of global codes is imcompible with the input dim of the network.")  # nerf_mlp.py:163-164
        lead = lengths.shape[:-1]
        P = lengths.shape[-1]
        if global_codes is not None:
            return self._forward_with_codes(origins, directions, lengths, global_codes)
        key = None
        if self._flat_leaf is not None:
            flat = self._flat_leaf
        elif torch.is_grad_enabled():
            flat = self._flat()
        else:  # inference: one flat copy per parameter version, not one concatenation per chunk
            key = self._param_key(lengths.device)
            if self._flat_nograd[0] != key:
                self._flat_nograd = [key, self._flat().detach()]
            flat = self._flat_nograd[1]
        need_grad = torch.is_grad_enabled() and flat.requires_grad
        plan = self.plan_for(flat, need_grad, key)
        o = N.f32c(origins.expand(*lead, 3)).reshape(-1, 3)
        d = N.f32c(directions.expand(*lead, 3)).reshape(-1, 3)
        z = N.f32c(lengths).reshape(-1, P)
        density, rgb = ops.MlpFunction.apply(flat, o, d, z, plan, need_grad,
                                             self._flat_grad if self._flat_leaf is not None else None)
        return dict(
            rays_densities=density.reshape(*lead, P, 1),
            rays_features=rgb.reshape(*lead, P, self.color_dim),
            aux={},
        )

    def _forward_with_codes(self, origins, directions, lengths, global_codes) -> dict:
        """latent_dim > 0: one kernel chain per image with that image's effective parameters (`_effective_flat`)."""
        lead, P = lengths.shape[:-1], lengths.shape[-1]
        B = lead[0]
        if global_codes.shape[0] != B:
            raise ValueError("The shape of global codes is imcompible with the input dim of the network.")
        o = N.f32c(origins.expand(*lead, 3)).reshape(B, -1, 3)
        d = N.f32c(directions.expand(*lead, 3)).reshape(B, -1, 3)
        z = N.f32c(lengths).reshape(B, -1, P)
        need_grad = torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters()) or global_codes.requires_grad)
        fmt = self._fmt[bool(need_grad)]
        dens, rgbs = [], []
        for b in range(B):
            flat = self._effective_flat(global_codes[b].to(o.dtype))
            # a private plan per call: the packed image belongs to this image's code (and autograd keeps it until backward)
            plan = ops.MlpPlan.create(N.MlpArch(*self._arch_fields, fmt), flat.device)
            plan.pack(flat.detach())
            density, rgb = ops.MlpFunction.apply(flat, o[b], d[b], z[b], plan, need_grad, None)
            dens.append(density)
            rgbs.append(rgb)
        return dict(
            rays_densities=torch.stack(dens).reshape(*lead, P, 1),
            rays_features=torch.stack(rgbs).reshape(*lead, P, self.color_dim),
            aux={},
        )
