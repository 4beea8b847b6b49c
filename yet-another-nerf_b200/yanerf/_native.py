"""ctypes binding of libyanerf_b200.so (include/yanerf_b200.h).

The product path has NO fallback: if the shared library is missing or a kernel
call fails, the operators raise.  Tensors cross the boundary as raw device
pointers + sizes; the stream is torch's current CUDA stream.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint32, c_void_p
from typing import Optional

import torch

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG_ROOT, "libyanerf_b200.so")

FMT_FP16 = 0
FMT_BF16 = 1


class MlpArch(Structure):
    """Mirror of `yn_mlp_arch`."""

    _fields_ = [
        ("n_layers", c_int32),
        ("skip_mask", c_uint32),
        ("n_freq_xyz", c_int32),
        ("n_freq_dir", c_int32),
        ("hidden_last", c_int32),
        ("hidden_dir", c_int32),
        ("color_dim", c_int32),
        ("fmt", c_int32),
    ]


class MarchCfg(Structure):
    """Mirror of `yn_march_cfg`."""

    _fields_ = [
        ("background_opacity", c_float),
        ("background_density_bias", c_float),
        ("density_noise_std", c_float),
        ("blend_output", c_int32),
        ("hard_background", c_int32),
        ("bg_channels", c_int32),
        ("bg_const", c_float * 4),
    ]


# name -> (restype, argtypes); every symbol include/yanerf_b200.h declares
_P = c_void_p
SYMBOLS = {
    "yn_version": (c_int, []),
    "yn_last_error_string": (c_char_p, []),
    "yn_ray_bundle": (c_int, [_P, c_int64, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int64, c_int, c_int, c_int, c_int, _P]),
    "yn_sample_pixels": (c_int, [_P, _P, _P, c_int64, c_int64, c_int, c_int, _P]),
    "yn_mlp_param_count": (c_int64, [POINTER(MlpArch)]),
    "yn_mlp_wpack_bytes": (c_int64, [POINTER(MlpArch)]),
    "yn_mlp_aux_floats": (c_int64, [POINTER(MlpArch)]),
    "yn_mlp_stash_bytes": (c_int64, [POINTER(MlpArch), c_int64]),
    "yn_mlp_pack_weights": (c_int, [POINTER(MlpArch), _P, _P, _P, _P]),
    "yn_mlp_dirbias": (c_int, [POINTER(MlpArch), _P, _P, _P, c_int64, _P]),
    "yn_mlp_fwd": (c_int, [POINTER(MlpArch), _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, _P]),
    "yn_mlp_bwd_workspace_bytes": (c_int64, [POINTER(MlpArch), c_int64]),
    "yn_mlp_bwd": (c_int, [POINTER(MlpArch), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, _P]),
    "yn_composite_fwd": (c_int, [POINTER(MarchCfg), _P, _P, _P, _P, _P, _P, c_int, _P, _P, _P, _P, _P, c_int64, c_int, c_int, _P]),
    "yn_composite_bwd": (c_int, [POINTER(MarchCfg), _P, _P, _P, _P, _P, _P, c_int, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, _P]),
    "yn_sample_pdf_merge": (c_int, [_P, _P, _P, c_int64, _P, c_int, _P, _P, _P, c_int64, c_int, c_int, c_int, _P]),
    "yn_step_begin": (c_int, [_P, _P, _P]),
    "yn_train_rays": (c_int, [_P, c_int, _P, c_int64, c_int64, _P, _P, c_int, _P, _P, _P, _P, _P, c_int64, c_int64, c_int, c_int, c_int, _P]),
    "yn_rgb_loss_scratch_bytes": (c_int64, [c_int64]),
    "yn_rgb_loss_fwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int64, c_int, c_int, c_int, _P]),
    "yn_rng_fill": (c_int, [_P, c_int, c_int, _P, c_int64, c_int, _P]),
    "yn_scatter_rays": (c_int, [_P, _P, _P, c_int, _P, c_int64, c_int64, c_int, c_int, _P]),
    "yn_rgb_loss_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int64, c_int64, c_int, c_int, c_int, _P]),
    "yn_sample_pdf": (c_int, [_P, _P, _P, c_int64, _P, _P, _P, c_int64, c_int, c_int, _P]),
    "yn_adam_step": (c_int, [_P, _P, _P, _P, c_int64, c_float, c_double, c_double, c_float, c_int32, c_float, _P]),
    "yn_adam_step_dev": (c_int, [_P, _P, _P, _P, c_int64, _P, c_double, c_double, c_float, c_float, _P]),
}

_lib: Optional[ctypes.CDLL] = None


def lib() -> ctypes.CDLL:
    """Load the shared library once; raise (never fall back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or torch fallback for the yanerf hot path)"
            )
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


_ERRORS = {-1: ValueError, -2: NotImplementedError, -3: RuntimeError, -4: RuntimeError}


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().yn_last_error_string().decode()
        raise _ERRORS.get(rc, RuntimeError)(msg)


def stream_ptr(device_index: Optional[int] = None) -> c_void_p:
    """Raw handle of torch's current stream on `device_index` (default: the current device)."""
    if device_index is None:
        device_index = torch.cuda.current_device()
    return c_void_p(torch._C._cuda_getCurrentRawStream(device_index))


def ptr(t: Optional[torch.Tensor], dtype=torch.float32) -> c_void_p:
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return c_void_p(0)
    if not t.is_cuda:
        raise RuntimeError("yanerf kernels need CUDA tensors (there is no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return c_void_p(t.data_ptr())


def f32c(t: torch.Tensor) -> torch.Tensor:
    """fp32 contiguous view/copy (no-op for the tensors the pipeline produces itself)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()
