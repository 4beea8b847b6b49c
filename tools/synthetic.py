"""Deterministic synthetic inputs shared by bench.py, the tests and the golden
generator.  numpy's legacy MT19937 `RandomState` stream is stable across numpy
versions, so every party regenerates identical weights / cameras / draws from a
seed instead of shipping multi-megabyte fixtures.

Shapes follow SURVEY §8(d): camera `pose = [I | (0,0,-4)]`, focal 1111.111 for
800x800 (lego) and 407.6 for 378x504 (fern shape).
"""
from __future__ import annotations

import math
from typing import Dict, Sequence, Tuple

import numpy as np
import torch


def _uniform(rs: np.random.RandomState, shape, bound: float) -> torch.Tensor:
    return torch.from_numpy(rs.uniform(-bound, bound, size=shape).astype(np.float32))


def synth_mlp_state(shapes: "Dict[str, Tuple[int, ...]]", seed: int, gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """Xavier-uniform-scaled weights (x `gain`) and small uniform biases for the
    NeRFMLP parameter table `shapes` (name -> shape), reproducible from `seed`.
    `gain` > 1 gives a 'trained-scale' net with non-trivial densities."""
    rs = np.random.RandomState(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shape in shapes.items():
        if name.endswith("weight"):
            fan_out, fan_in = shape
            out[name] = _uniform(rs, shape, gain * math.sqrt(6.0 / (fan_in + fan_out)))
        else:
            out[name] = _uniform(rs, shape, 0.05)
    return out


def synth_camera(batch: int, seed: int = 0, jitter: float = 0.05) -> torch.Tensor:
    """Camera-to-world [B,3,4]: identity rotation perturbed by a small rotation,
    centre near (0,0,-4) looking down +z (depth range [2,6] covers the origin)."""
    rs = np.random.RandomState(1000 + seed)
    poses = []
    for _ in range(batch):
        a = rs.uniform(-jitter, jitter, size=3)
        rx = np.array([[1, 0, 0], [0, math.cos(a[0]), -math.sin(a[0])], [0, math.sin(a[0]), math.cos(a[0])]])
        ry = np.array([[math.cos(a[1]), 0, math.sin(a[1])], [0, 1, 0], [-math.sin(a[1]), 0, math.cos(a[1])]])
        rz = np.array([[math.cos(a[2]), -math.sin(a[2]), 0], [math.sin(a[2]), math.cos(a[2]), 0], [0, 0, 1]])
        rot = rz @ ry @ rx
        t = np.array([0.0, 0.0, -4.0]) + rs.uniform(-jitter, jitter, size=3)
        poses.append(np.concatenate([rot, t[:, None]], axis=1))
    return torch.from_numpy(np.stack(poses).astype(np.float32))


def synth_image(batch: int, height: int, width: int, seed: int = 1) -> torch.Tensor:
    rs = np.random.RandomState(2000 + seed)
    return torch.from_numpy(rs.uniform(0, 1, size=(batch, height, width, 3)).astype(np.float32))


def synth_draws(batch: int, n_rays: int, n_pixels: int, n_coarse: int, n_fine: int, seed: int = 0) -> Dict[str, torch.Tensor]:
    """The five random draws one training forward consumes, in the reference's
    call order: multinomial pixel pick (without replacement),
    stratified-jitter uniforms, coarse density noise, sample_pdf uniforms, fine
    density noise."""
    rs = np.random.RandomState(3000 + seed)
    pix = np.stack([rs.choice(n_pixels, size=n_rays, replace=False) for _ in range(batch)]).astype(np.int64)
    R = batch * n_rays
    f32 = np.float32
    # uniforms strictly inside [0,1) like torch.rand
    u_strat = np.minimum(rs.uniform(0, 1, size=(batch, n_rays, n_coarse)).astype(f32), f32(1 - 2**-24))
    noise0 = rs.standard_normal(size=(R, n_coarse)).astype(f32)
    u_pdf = np.minimum(rs.uniform(0, 1, size=(R, n_fine)).astype(f32), f32(1 - 2**-24))
    noise1 = rs.standard_normal(size=(R, n_coarse + n_fine)).astype(f32)
    return {
        "pix": torch.from_numpy(pix),
        "u_strat": torch.from_numpy(u_strat),
        "noise0": torch.from_numpy(noise0),
        "u_pdf": torch.from_numpy(u_pdf),
        "noise1": torch.from_numpy(noise1),
    }


LEGO_FOCAL = 1111.111
FERN_FOCAL = 407.6


def orbit_cameras(n: int, radius: float = 4.0, elevation: float = 0.15) -> torch.Tensor:
    """`n` camera-to-world poses [n,3,4] on a circle around the origin, looking at it (x right, y down, z forward:
    the convention of the reference's README.md:99-103)."""
    poses = []
    for i in range(n):
        th = 2.0 * math.pi * i / n
        c = np.array([radius * math.sin(th), -radius * math.sin(elevation), -radius * math.cos(th)])
        fwd = -c / np.linalg.norm(c)
        right = np.cross(np.array([0.0, 1.0, 0.0]), fwd)
        right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        poses.append(np.concatenate([np.stack([right, down, fwd], axis=1), c[:, None]], axis=1))
    return torch.from_numpy(np.stack(poses).astype(np.float32))


def sphere_scene_images(poses: torch.Tensor, focal: float, height: int, width: int, radius: float = 1.0) -> torch.Tensor:
    """Analytic multi-view ground truth [n,H,W,3]: a unit sphere at the origin whose colour is 0.5 + 0.5 * normal, on a
    black background, seen through the pinhole model of `_xy_to_ray_bundle` (ray_sampler.py:297-312)."""
    ys, xs = torch.meshgrid(torch.arange(height, dtype=torch.float32), torch.arange(width, dtype=torch.float32), indexing="ij")
    dcam = torch.stack(((xs - width / 2) / focal, (ys - height / 2) / focal, torch.ones_like(xs)), dim=-1)  # [H,W,3]
    out = []
    for pose in poses:
        rot, o = pose[:, :3], pose[:, 3]
        d = dcam @ rot.T
        a = (d * d).sum(-1)
        b = 2.0 * (d * o).sum(-1)
        c = float((o * o).sum()) - radius * radius
        disc = b * b - 4 * a * c
        hit = disc > 0
        t = (-b - torch.sqrt(disc.clamp_min(0.0))) / (2 * a)
        normal = (o + t[..., None] * d) / radius
        out.append(torch.where(hit[..., None], 0.5 + 0.5 * normal, torch.zeros(3)))
    return torch.stack(out).float()
