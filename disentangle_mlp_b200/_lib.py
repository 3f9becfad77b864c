"""ctypes binding of libdm_b200.so (see include/dm_b200.h).

There is no fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

from .build import LIB_PATH

c_void_p, c_int, c_float, c_ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong


class BnFuse(C.Structure):
    """dm_bn_fuse: BatchNorm statistics + finalize fused into the producing kernel (scratch NULL = off)."""
    _fields_ = [("scratch", c_void_p), ("groups", c_int), ("rows", c_ll), ("gamma", c_void_p), ("beta", c_void_p),
                ("running_mean", c_void_p), ("running_var", c_void_p), ("num_batches_tracked", c_void_p),
                ("momentum", c_float), ("eps", c_float), ("scale_shift", c_void_p), ("mean_invstd", c_void_p)]


class GemmDesc(C.Structure):
    _fields_ = [
        ("layout", c_int), ("m", c_int), ("n", c_int), ("k", c_int),
        ("a", c_void_p), ("lda", c_ll), ("b", c_void_p), ("ldb", c_ll),
        ("d", c_void_p), ("ldd_m", c_ll), ("ldd_n", c_ll),
        ("d_f32", c_int), ("accumulate", c_int), ("bias", c_void_p),
        ("m_store", c_int), ("n_store", c_int), ("splits", c_int), ("k_alg", c_int),
        ("bn", BnFuse),
    ]


class ConvGeom(C.Structure):
    _fields_ = [("batch", c_int), ("hs", c_int), ("ws", c_int), ("cs", c_int),
                ("hb", c_int), ("wb", c_int), ("cb", c_int), ("stride", c_int)]


GEMM_NT, GEMM_NN, GEMM_TN = 0, 1, 2

# name -> argtypes (all return int except the first three)
_SIGS = {
    "dm_gemm_bf16": [C.POINTER(GemmDesc), c_void_p],
    "dm_conv_down": [C.POINTER(ConvGeom), c_void_p, c_void_p, c_void_p, c_void_p, C.POINTER(BnFuse), c_void_p],
    "dm_conv_up": [C.POINTER(ConvGeom), c_void_p, c_void_p, c_void_p, c_void_p, c_int, C.POINTER(BnFuse), c_void_p],
    "dm_conv_wgrad": [C.POINTER(ConvGeom), c_void_p, c_void_p, c_void_p, c_int, c_void_p],
    "dm_gemm_tf32": [C.POINTER(GemmDesc), c_void_p],
    "dm_conv_down_tf32": [C.POINTER(ConvGeom), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "dm_conv_up_tf32": [C.POINTER(ConvGeom), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "dm_conv_wgrad_tf32": [C.POINTER(ConvGeom), c_void_p, c_void_p, c_void_p, c_void_p],
    "dm_unpack_conv_grad": [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p],
    "dm_profile_enable": [c_int],
    "dm_profile_dump": [C.c_char_p],
    "dm_profile_read": [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(c_ll)],
    "dm_debug_last_plan": [C.POINTER(c_int), C.POINTER(c_int), C.POINTER(c_int)],
    "dm_bn_parts": [c_ll, c_int],
    "dm_bn_slots": [],
    "dm_bn_stats": [c_void_p, c_int, c_int, C.POINTER(BnFuse), c_void_p],
    "dm_bn_apply_act": [c_void_p, c_int, c_ll, c_int, c_void_p, c_int, c_float, c_void_p, c_int, c_void_p],
    "dm_bn_forward": [c_void_p, c_int, c_ll, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float,
                      c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p],
    "dm_linear_pair_forward": [c_void_p] * 6 + [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p],
    "dm_linear_pair_backward": [c_void_p] * 6 + [c_int, c_int, c_int] + [c_void_p] * 7,
    "dm_bn1d_forward": [c_void_p, c_ll, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                        c_float, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p],
    "dm_bn1d_backward": [c_void_p, c_void_p, c_ll, c_int, c_int, c_void_p, c_void_p, c_int, c_float, c_void_p, c_ll,
                         c_void_p, c_void_p, c_void_p],
    "dm_bn_backward": [c_void_p, c_void_p, c_int, c_ll, c_int, c_void_p, c_void_p, c_int, c_float, c_void_p,
                       c_void_p, c_void_p, c_void_p, c_int, c_void_p],
    "dm_bias_act": [c_void_p, c_ll, c_int, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p],
    "dm_act_backward": [c_void_p, c_void_p, c_ll, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p],
    "dm_colsum": [c_void_p, c_int, c_ll, c_int, c_void_p, c_void_p, c_void_p],
    "dm_im2col3": [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p],
    "dm_nhwc3_to_nchw": [c_void_p, c_ll, c_int, c_int, c_void_p, c_void_p, c_void_p],
    "dm_tanh_backward": [c_void_p, c_void_p, c_ll, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    "dm_pad_image3": [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p],
    "dm_pack_conv3_weights": [c_void_p, c_int, c_int, c_void_p, c_void_p],
    "dm_conv3_fwd": [C.POINTER(ConvGeom), c_void_p, c_void_p, c_void_p, c_void_p, C.POINTER(BnFuse), c_void_p],
    "dm_conv3_wgrad": [C.POINTER(ConvGeom), c_void_p, c_void_p, c_void_p, c_void_p],
    "dm_unpack_conv3_grad": [c_void_p, c_int, c_int, c_void_p, c_void_p],
    "dm_transpose_bf16": [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p],
    "dm_pack_conv_weights": [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    "dm_pack_up_merged": [c_void_p, c_int, c_int, c_void_p, c_void_p],
    "dm_pack_down_pairs": [c_void_p, c_int, c_int, c_void_p, c_void_p],
    "dm_conv_down_paired": [C.POINTER(ConvGeom), c_void_p, c_void_p, c_void_p, c_void_p, C.POINTER(BnFuse), c_void_p],
    "dm_conv_up_merged": [C.POINTER(ConvGeom), c_void_p, c_void_p, c_void_p, c_void_p, C.POINTER(BnFuse), c_void_p],
    "dm_cast_bf16": [c_void_p, c_ll, c_void_p, c_void_p],
    "dm_reparam_forward": [c_void_p, c_void_p, c_void_p, c_ll, c_void_p, c_void_p, c_void_p],
    "dm_reparam_backward": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_void_p, c_void_p, c_void_p,
                            c_void_p, c_void_p],
    "dm_head_forward": [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    "dm_head_backward": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                         c_void_p, c_void_p],
    "dm_mse_sum": [c_void_p, c_void_p, c_ll, c_float, c_void_p, c_float, c_int, c_void_p, c_void_p],
    "dm_kl": [c_void_p, c_void_p, c_ll, c_float, c_void_p, c_int, c_void_p, c_void_p, c_void_p],
    "dm_bce_const": [c_void_p, c_int, c_float, c_float, c_void_p, c_float, c_void_p, c_int, c_void_p, c_void_p, c_void_p],
    "dm_adam_step": [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, C.c_double, C.c_double, C.c_double, C.c_double, c_int,
                     c_void_p, c_float, c_void_p, c_void_p],
    "dm_adam_step_gated": [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_ll, C.c_double, C.c_double, C.c_double,
                           C.c_double, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p],
    "dm_adam_step_ex": [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_ll, C.c_double, C.c_double, C.c_double,
                        C.c_double, c_int, c_void_p, c_int, c_float, c_void_p, c_void_p],
}

#: every symbol include/dm_b200.h declares
EXPORTED = ["dm_last_error", "dm_version", "dm_launch_count", "dm_bn_scratch_floats", "dm_pim_elems", "dm_workspace_bytes",
            *list(_SIGS)]

_lib = None


def load():
    """Load libdm_b200.so, (re)building it first when it is missing or its sources changed (content hash; the
    build is serialised by a file lock, so concurrent ranks do not race).  Raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    if os.environ.get("DM_B200_LIB"):
        path = Path(os.environ["DM_B200_LIB"])
    else:
        from .build import build

        path = build(force=bool(os.environ.get("DM_B200_REBUILD")))
    lib = C.CDLL(str(path))
    lib.dm_last_error.restype = C.c_char_p
    lib.dm_last_error.argtypes = []
    lib.dm_version.restype = c_int
    lib.dm_launch_count.restype = c_ll
    lib.dm_bn_scratch_floats.restype = c_ll
    lib.dm_bn_scratch_floats.argtypes = [c_int, c_int]
    lib.dm_pim_elems.restype = c_ll
    lib.dm_pim_elems.argtypes = [c_int]
    lib.dm_workspace_bytes.restype = c_ll
    lib.dm_workspace_bytes.argtypes = [c_int, C.POINTER(c_ll), c_int]
    for name, args in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = c_int
    _lib = lib
    return lib


_DEBUG_SYNC = bool(os.environ.get("DM_DEBUG_SYNC"))


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().dm_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")
    if _DEBUG_SYNC:  # debugging aid: surface asynchronous faults at the call that caused them
        import torch

        try:
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            raise RuntimeError(f"{what}: device fault after launch: {e}") from e


def launch_count() -> int:
    return int(load().dm_launch_count())
