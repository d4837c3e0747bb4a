"""ctypes binding of libuspace_b200.so (the C ABI declared in include/uspace_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C uspace_b200/csrc``.  There is no
fallback: if the shared object is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libuspace_b200.so")

USP_OK = 0
METHOD = {"euler": 0, "heun": 1, "midpoint": 2, "rk4": 3}
ADAPTIVE_METHOD = {"dopri5": 4, "bosh3": 5, "adaptive_heun": 6}
EDIT_LOC = {None: 0, "none": 0, "head": 1, "tail": 2}
OPERAND = {"bf16": 0, "fp16": 1}
EPI = {"qkv": 0, "bias_gelu": 1, "bias_resid": 2, "bias_f32": 3}


class UspConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "img_size", "patch_size", "in_chans", "embed_dim", "depth", "num_heads", "mlp_hidden", "num_classes",
        "clip_dim", "num_clip_token", "qkv_bias", "conv", "skip", "operand_dtype", "fuse_layernorm", "mlp_time_embed")]


class UspAttnEdit(C.Structure):
    _fields_ = [("colscale", C.c_void_p), ("block_mask", C.c_uint64), ("t_edit", C.c_float)]


class UspAdaptiveStats(C.Structure):
    _fields_ = [("n_accept", C.c_int32), ("n_reject", C.c_int32), ("nfe", C.c_int32), ("last_ratio", C.c_float),
                ("last_dt", C.c_double)]


_lib = None

_vp, _i, _f, _i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64
_SIGNATURES = {
    "usp_create": (_i, [C.POINTER(UspConfig), _i, C.POINTER(_vp)]),
    "usp_destroy": (None, [_vp]),
    "usp_last_error": (C.c_char_p, [_vp]),
    "usp_set_weight": (_i, [_vp, C.c_char_p, _vp, C.POINTER(_i64), _i]),
    "usp_finalize_weights": (_i, [_vp, _vp]),
    "usp_num_weights": (_i, [_vp]),
    "usp_weight_name": (C.c_char_p, [_vp, _i]),
    "usp_forward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "usp_forward_edit": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, C.POINTER(UspAttnEdit), _vp]),
    "usp_forward_hook": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _f, _vp, C.POINTER(UspAttnEdit), _vp]),
    "usp_sample_edit": (_i, [_vp, _vp, _vp, _vp, _i, _f, _f, _f, _i, _vp, _f, _f, _i, C.POINTER(UspAttnEdit), _vp]),
    "usp_sample": (_i, [_vp, _vp, _vp, _vp, _i, _f, _f, _f, _i, _vp, _f, _f, _i, _vp]),
    "usp_sample_adaptive": (_i, [_vp, _vp, _vp, _vp, _i, _f, _f, _i, C.c_double, C.c_double, _vp, _i, _f, _f, _i,
                                 C.POINTER(UspAttnEdit), _i, C.POINTER(UspAdaptiveStats), _vp]),
    "usp_sample_adaptive_read": (_i, [_vp, _vp, _vp, _vp, _i, _f, _f, _i, C.c_double, C.c_double, _i, _vp, _vp, _i,
                                      C.POINTER(_i), _i, C.POINTER(UspAdaptiveStats), _vp]),
    "usp_sample_sweep": (_i, [_vp, _vp, _vp, _vp, _vp, _i, C.POINTER(_f), _i, _f, _f, _f, _i, _vp, _f, _i, _vp]),
    "usp_sample_read": (_i, [_vp, _vp, _vp, _vp, _i, _f, _f, _f, _i, _i, _vp, _vp]),
    "usp_sample_host": (_i, [_vp, _vp, _vp, _vp, _i, _f, _f, _f, _i, _vp, _f, _f, _i]),
    "usp_nonfinite": (_i, [_vp, C.POINTER(_i), _vp]),
    "usp_grid_size": (_i, [_f, _f, _f]),
    "usp_time_grid": (_i, [_f, _f, _f, C.POINTER(_f), _i]),
    "usp_workspace_bytes": (C.c_size_t, [_vp, _i]),
    "usp_kernels_per_forward": (_i, [_vp]),
    "usp_flops_per_forward": (C.c_double, [_vp]),
    "usp_last_forward_ms": (_i, [_vp, C.POINTER(_f)]),
    "usp_profile_forward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, C.POINTER(_f), C.POINTER(_i), _vp]),
    "usp_profile_forward_n": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, C.POINTER(_f), C.POINTER(_i), _vp]),
    "usp_vae_create": (_i, [_i, _f, C.POINTER(_vp)]),
    "usp_vae_destroy": (None, [_vp]),
    "usp_vae_last_error": (C.c_char_p, [_vp]),
    "usp_vae_num_weights": (_i, [_vp]),
    "usp_vae_weight_name": (C.c_char_p, [_vp, _i]),
    "usp_vae_set_weight": (_i, [_vp, C.c_char_p, _vp, C.POINTER(_i64), _i]),
    "usp_vae_set_precision": (_i, [_vp, _i]),
    "usp_vae_finalize": (_i, [_vp, _vp]),
    "usp_vae_workspace_bytes": (C.c_size_t, [_vp]),
    "usp_vae_release_workspace": (_i, [_vp]),
    "usp_vae_decode": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "usp_vae_encode_moments": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "usp_op_convert16": (_i, [_vp, _vp, _i64, _i, _vp]),
    "usp_op_gemm": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "usp_op_attention": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "usp_op_layernorm": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "usp_op_patch_embed": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "usp_op_unpatchify_conv": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)
KERNEL_CLASSES = ("embed", "layernorm", "gemm_qkv", "attention", "gemm_proj", "gemm_fc1", "gemm_fc2", "gemm_skip",
                  "head_final", "context_embed")


def load():
    """Load (once) and return the ctypes library; raises if the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(uspace_b200 has no CPU / PyTorch fallback for the sampling path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, handle=None, what: str = "uspace_b200", vae: bool = False):
    if rc != USP_OK:
        msg = load().usp_vae_last_error(handle) if (vae or what.startswith("usp_vae")) else load().usp_last_error(handle)
        raise RuntimeError(f"{what} failed (status {rc}): {msg.decode() if msg else '?'}")
