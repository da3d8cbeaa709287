"""ctypes binding of ``libmtv_b200.so`` (C ABI in ``include/mtv_b200.h``).

This is the binding a MoDiTalker maintainer would add.  There
is deliberately NO fallback: if the shared library is missing or a call fails the
caller gets a ``RuntimeError`` — the hot path never silently routes through
PyTorch or the CPU.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, byref, c_char, c_char_p, c_float, c_int32, c_int64, c_void_p
from typing import Optional

MTV_MAX_LEVELS = 8
MTV_ABI_VERSION = 1

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmtv_b200.so")


class MtvConfig(Structure):
    _fields_ = [
        ("abi_version", c_int32),
        ("image_size", c_int32),
        ("in_channels", c_int32),
        ("out_channels", c_int32),
        ("model_channels", c_int32),
        ("num_res_blocks", c_int32),
        ("num_heads", c_int32),
        ("num_levels", c_int32),
        ("channel_mult", c_int32 * MTV_MAX_LEVELS),
        ("attn_at_level", c_int32 * MTV_MAX_LEVELS),
        ("device", c_int32),
        ("kernel_path", c_int32),
    ]


class MtvKernelTime(Structure):
    _fields_ = [("name", c_char * 48), ("us", c_float), ("flops", c_float), ("bytes", c_float)]


_SIGNATURES = {
    # name: (restype, argtypes)
    "mtv_abi_version": (c_int32, []),
    "mtv_last_error": (c_char_p, []),
    "mtv_create": (c_int32, [POINTER(MtvConfig), POINTER(c_void_p)]),
    "mtv_destroy": (c_int32, [c_void_p]),
    "mtv_load_weight": (c_int32, [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int32, POINTER(c_int32), c_void_p]),
    "mtv_weights_ready": (c_int32, [c_void_p, POINTER(c_int32), POINTER(c_int32)]),
    "mtv_num_weight_names": (c_int32, [c_void_p]),
    "mtv_weight_name": (c_char_p, [c_void_p, c_int32]),
    "mtv_unet_forward": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int32, c_void_p, c_void_p]),
    "mtv_ddim_step": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float,
                                c_float, c_int32, c_void_p]),
    "mtv_q_sample": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_void_p, c_void_p]),
    "mtv_io_prep_frames": (c_int32, [c_int32, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_int32, c_void_p, c_void_p]),
    "mtv_io_prep_frames_ex": (c_int32, [c_int32, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "mtv_io_rasterize_landmarks": (c_int32, [c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p,
                                             c_void_p]),
    "mtv_io_frames_out": (c_int32, [c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_int32,
                                    c_void_p]),
    "mtv_plan_info": (c_int32, [c_void_p, c_int32, POINTER(c_int64), POINTER(c_int64), POINTER(c_int64)]),
    "mtv_debug_read": (c_int32, [c_void_p, c_char_p, c_void_p, c_int64, c_void_p]),
    "mtv_debug_tc_timing": (c_int32, [c_void_p, c_void_p, c_int32, POINTER(c_int32)]),
    "mtv_profile_forward": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int32, c_void_p,
                                      POINTER(MtvKernelTime), c_int32, POINTER(c_int32), c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib: Optional[ctypes.CDLL] = None


def load_library(path: Optional[str] = None) -> ctypes.CDLL:
    """dlopen the library and attach signatures.  Raises RuntimeError when the
    shared object is absent (run ``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("MTV_B200_LIB", LIB_PATH)
    if not os.path.exists(p):
        raise RuntimeError(
            f"moditalker_b200: CUDA library not found at {p}. Build it with __graft_entry__.build(); "
            "there is no CPU / PyTorch fallback for the denoising hot path."
        )
    lib = ctypes.CDLL(p)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here == missing export
        fn.restype = res
        fn.argtypes = args
    if lib.mtv_abi_version() != MTV_ABI_VERSION:
        raise RuntimeError("moditalker_b200: libmtv_b200.so ABI version mismatch; rebuild")
    if path is None:
        _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load_library().mtv_last_error()
        raise RuntimeError(f"{what}: {msg.decode() if msg else 'unknown error'}")
