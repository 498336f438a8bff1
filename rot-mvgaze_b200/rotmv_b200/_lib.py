"""ctypes binding of librotmv_sm100.so (the C ABI declared in include/rotmv_sm100.h).

There is no fallback: if the shared library is missing, or a call returns non-zero, this module
raises. PyTorch is used only for device memory and streams; every FLOP of the product path runs in
the hand-written sm_100a kernels behind these entry points.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

F32 = 0
BF16 = 1
ENGINE_AUTO = 0
ENGINE_SIMT = 1
ENGINE_TC = 2

_LIB_NAME = "librotmv_sm100.so"
_lib = None


class RotmvError(RuntimeError):
    pass


class ConvArgs(C.Structure):
    """Mirror of `struct rmv_conv_args` (include/rotmv_sm100.h)."""

    _fields_ = [
        ("x_dtype", C.c_int), ("y_dtype", C.c_int), ("engine", C.c_int), ("block_n", C.c_int),
        ("x", C.c_void_p),
        ("x_sn", C.c_longlong), ("x_sh", C.c_longlong), ("x_sw", C.c_longlong), ("x_sc", C.c_longlong),
        ("n_img", C.c_int), ("in_h", C.c_int), ("in_w", C.c_int), ("c_in", C.c_int),
        ("w", C.c_void_p),
        ("c_out", C.c_int), ("kh", C.c_int), ("kw", C.c_int), ("stride", C.c_int), ("pad", C.c_int),
        ("y", C.c_void_p),
        ("y_sn", C.c_longlong), ("y_sh", C.c_longlong), ("y_sw", C.c_longlong),
        ("out_h", C.c_int), ("out_w", C.c_int),
        ("scale", C.c_void_p), ("shift", C.c_void_p), ("residual", C.c_void_p),
        ("r_sn", C.c_longlong), ("r_sh", C.c_longlong), ("r_sw", C.c_longlong),
        ("relu", C.c_int),
        ("stat_acc", C.c_void_p), ("stat_views", C.c_int), ("stat_finalize", C.c_void_p),
        ("bn_mode", C.c_int), ("bn_a", C.c_void_p), ("bn_b", C.c_void_p), ("bn_c", C.c_void_p),
        ("bn_bits", C.c_void_p), ("mask_bits", C.c_void_p), ("mask_off", C.c_longlong),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class BnParams(C.Structure):
    """Mirror of `struct rmv_bn_params`."""

    _fields_ = [("ticket", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("running_mean", C.c_void_p), ("running_var", C.c_void_p), ("num_batches", C.c_void_p),
                ("mean", C.c_void_p), ("invstd", C.c_void_p), ("a", C.c_void_p), ("b", C.c_void_p),
                ("dgamma", C.c_void_p), ("dbeta", C.c_void_p), ("k0", C.c_void_p), ("k1", C.c_void_p),
                ("k2", C.c_void_p), ("eps", C.c_float), ("momentum", C.c_float)]


class PermuteJob(C.Structure):
    """Mirror of `struct rmv_permute_job`."""

    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p),
                ("d0", C.c_int), ("d1", C.c_int), ("d2", C.c_int), ("d3", C.c_int),
                ("s0", C.c_longlong), ("s1", C.c_longlong), ("s2", C.c_longlong), ("s3", C.c_longlong),
                ("flip1", C.c_int), ("flip2", C.c_int), ("dst_dtype", C.c_int),
                ("first_block", C.c_uint), ("kind", C.c_int)]


# name -> (restype, argtypes); every symbol include/rotmv_sm100.h declares must be listed here
# (tests/test_abi.py checks the header against this table and against the built library).
_vp, _i, _ll, _f = C.c_void_p, C.c_int, C.c_longlong, C.c_float
SIGNATURES = {
    "rmv_version": (_i, []),
    "rmv_set_tuning": (_i, [C.c_char_p, _i]),
    "rmv_last_error": (C.c_char_p, []),
    "rmv_device_check": (_i, [_i]),
    "rmv_conv2d_dgrad": (_i, [C.POINTER(ConvArgs), _vp]),
    "rmv_conv2d_fwd": (_i, [C.POINTER(ConvArgs), _vp]),
    "rmv_conv_bn_stats": (_i, [C.POINTER(ConvArgs), _vp, C.POINTER(BnParams), _vp]),
    "rmv_conv_bn_bwd_reduce": (_i, [C.POINTER(ConvArgs), _vp, _vp, _vp, _vp, C.POINTER(BnParams), _vp]),
    "rmv_stem_im2col": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "rmv_stem_pack_weights": (_i, [_vp, _vp, _vp]),
    "rmv_stem_conv_fwd_u8": (_i, [_vp, C.POINTER(C.c_float), C.POINTER(C.c_float), _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "rmv_stem_conv_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "rmv_stem_wgrad": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "rmv_stem_wgrad_workspace_bytes": (C.c_size_t, []),
    "rmv_splitk_workspace_bytes": (C.c_size_t, []),
    "rmv_bn_workspace_bytes": (C.c_size_t, [_i, _i]),
    "rmv_conv2d_wgrad_tc_workspace_bytes": (C.c_size_t, [C.POINTER(ConvArgs)]),
    "rmv_nchw_to_nhwc": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "rmv_maxpool3x3s2_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "rmv_avgpool_fwd": (_i, [_vp, _i, _i, _i, _i, _vp, _ll, _vp, _ll, _vp]),
    "rmv_rotate_gather_fwd": (_i, [_vp, _ll, _vp, _vp, _ll, _i, _i, _i, _i, _i, _vp]),
    "rmv_head_loss_fwd": (_i, [_vp, _ll, _i, _vp, _vp, _i, _i, _vp, _vp, _f, _i, _f, _vp, _vp]),
    "rmv_angular_error_accum": (_i, [_vp, _ll, _vp, _ll, _i, _vp, _vp]),
    "rmv_pose_to_rotations": (_i, [_vp, _vp, _i, _i, _vp]),
    "rmv_relative_rotations": (_i, [_vp, _vp, _i, _i, _vp]),
    # training step
    "rmv_bn_stats": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "rmv_bn_finalize": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _ll, _f, _f, _vp]),
    "rmv_bn_stats_finalize": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                   _vp, _f, _f, _vp]),
    "rmv_bn_bwd_reduce_finalize": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp,
                                        _vp, _vp, _vp, _vp, _vp, _vp]),
    "rmv_bn_apply": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "rmv_bn_bwd_reduce": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "rmv_bn_bwd_finalize": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _ll, _vp]),
    "rmv_bn_bwd_apply": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "rmv_relu_bwd": (_i, [_vp, _ll, _vp, _ll, _vp, _ll, _vp, _ll, _i, _i, _i, _vp]),
    "rmv_colsum": (_i, [_vp, _ll, _i, _i, _i, _vp, _vp]),
    "rmv_permute_cast": (_i, [_vp, _vp, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _i, _i, _i, _vp]),
    "rmv_permute_cast_batch": (_i, [_vp, _i, C.c_uint, _vp]),
    "rmv_dilate2": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "rmv_maxpool3x3s2_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "rmv_maxpool3x3s2_fwd_idx": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "rmv_maxpool3x3s2_bwd_idx": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "rmv_mask_bits": (_i, [_vp, _vp, _vp, _ll, _i, _vp]),
    "rmv_avgpool_bwd": (_i, [_vp, _ll, _vp, _i, _i, _i, _i, _vp]),
    "rmv_head_loss_bwd": (_i, [_vp, _vp, _vp, _ll, _i, _vp, _i, _i, _f, _i, _f, _vp, _ll, _vp, _vp,
                               _vp, _vp]),
    "rmv_conv2d_wgrad": (_i, [C.POINTER(ConvArgs), _vp, _vp, _vp]),
    "rmv_conv2d_wgrad_tc": (_i, [C.POINTER(ConvArgs), _vp, _vp, _vp]),
    "rmv_adam_step": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _i, _f, _vp]),
    # constructor variants: re-layout kernels
    "rmv_strided_copy": (_i, [_vp, _i, _ll, _ll, _ll, _vp, _i, _ll, _ll, _ll, _i, _i, _i, _vp, _i, _vp]),
    "rmv_intensity_bn_train": (_i, [_vp, _ll, _i, _i, _i, _vp, _f, _f, _vp, _vp]),
    "rmv_fill_zero": (_i, [_vp, C.c_size_t, _vp]),
}


def lib_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)


def load():
    """Load the shared library (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RotmvError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C rot-mvgaze_b200/csrc`). There is no CPU or PyTorch fallback."
        )
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


# Number of successful kernel-launching C-ABI calls since import (bench.py's `gpu_launches`).
STATS = {"launches": 0}


def check(rc: int, what: str) -> None:
    if rc == 0:
        if what != "rmv_device_check":
            STATS["launches"] += 1
        return
    if rc != 0:
        msg = load().rmv_last_error()
        raise RotmvError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def dtype_code(t: torch.dtype) -> int:
    if t == torch.float32:
        return F32
    if t == torch.bfloat16:
        return BF16
    raise RotmvError(f"unsupported dtype {t}")


def ptr(t):
    return None if t is None else t.data_ptr()
