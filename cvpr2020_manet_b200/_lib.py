"""ctypes binding of libmanet_b200.so (the C ABI declared in include/manet_b200.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is
raised.  PyTorch is used by the callers of this module only for device memory and streams.
"""
from __future__ import annotations

import ctypes
import os
import re
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint32, c_void_p

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "lib", "libmanet_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(PKG_DIR), "include", "manet_b200.h")

GM_NORMALIZE = 1
GM_DROP_UNLAB = 2
GM_ENGINE_SIMT = 4
GM_ENGINE_EXACT3 = 8
GM_REUSE_REF = 128
GM_ENGINE_FR = 512
LM_ENGINE_SIMT = 1
LM_ENGINE_TENSOR = 2
STEP_SERIAL = 16
STEP_STREAM = 32
STEP_STREAM_RESET = 64
STEP_NO_REF_CACHE = 256
DT_F32, DT_F16, DT_F64 = 0, 1, 2

_I64, _I, _P, _SZ, _F = c_int64, c_int, c_void_p, c_size_t, c_float

# name -> (restype, argtypes); must list every function include/manet_b200.h declares
SIGNATURES = {
    "manet_abi_version": (c_int, []),
    "manet_last_error": (c_char_p, []),
    "manet_check_device": (c_int, []),
    "manet_global_match_workspace_bytes": (_SZ, [_I64, _I64, _I, _I, _I]),
    "manet_global_match": (c_int, [_P, _I64, _I64, _I64, _P, _P, _I64, _I64, _I64, _I, _I, _I, c_uint32, _P, _P, _P, _SZ, _P]),
    "manet_global_match_masked": (c_int, [_P, _I64, _I64, _I64, _P, _P, _I64, _I64, _I64, _I, _I, _I, _P, _P, _SZ, _P]),
    "manet_global_match_topk": (c_int, [_P, _I64, _I64, _I64, _P, _P, _I64, _I64, _I64, _I, _I, _I, _P, _P]),
    "manet_pairwise_sqdist": (c_int, [_P, _I64, _I64, _I64, _P, _I64, _I64, _I64, _I, _P, _P, _P, _P]),
    "manet_row_sqnorm": (c_int, [_P, _I64, _I64, _I64, _I, _P, _P]),
    "manet_select_labelled_workspace_bytes": (_SZ, [_I64]),
    "manet_select_labelled": (c_int, [_P, _I64, _P, _I64, _I64, _I, _P, _P, _P, _P, _SZ, _P]),
    "manet_local_match_workspace_bytes": (_SZ, [_I, _I, _I, _I, _I]),
    "manet_local_match": (c_int, [_P, _I64, _I64, _I64, _P, _I64, _I64, _I64, _P, _P, _I, _I, _I, _I, _I, _P, _P, _SZ, _P]),
    "manet_local_match_guard_stats": (c_int, [_P, _SZ, _I, _I, _I, _I, _I, POINTER(c_float), _P]),
    "manet_local_match_ex": (c_int, [_P, _I64, _I64, _I64, _P, _I64, _I64, _I64, _P, _P, _I, _I, _I, _I, _I, c_uint32, _P, _P, _SZ, _P]),
    "manet_local_window_distances_ex": (c_int, [_P, _I64, _I64, _I64, _P, _I64, _I64, _I64, _I, _I, _I, _I, c_uint32, _P, _P, _SZ, _P]),
    "manet_local_window_distances": (c_int, [_P, _I64, _I64, _I64, _P, _I64, _I64, _I64, _I, _I, _I, _I, _P, _P, _SZ, _P]),
    "manet_global_match_argmin": (c_int, [_P, _I64, _I64, _I64, _P, _P, _I64, _I64, _I64, _I, _I, _P, _P, _P]),
    "manet_global_match_argmin_ws": (c_int, [_P, _I64, _I64, _I64, _P, _P, _I64, _I64, _I64, _I, _I, c_uint32, _P, _P, _P, _SZ, _P]),
    "manet_global_match_backward": (c_int, [_P, _I64, _I64, _I64, _P, _I64, _I64, _I64, _I, _I, _P, _P, _P, _P, _P]),
    "manet_local_match_grad_workspace_bytes": (_SZ, [_I, _I, _I, _I, _I]),
    "manet_local_match_argmin": (c_int, [_P, _I64, _I64, _I64, _P, _I64, _I64, _I64, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _SZ, _P]),
    "manet_local_match_backward": (c_int, [_P, _I64, _I64, _I64, _P, _I64, _I64, _I64, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _SZ, _P]),
    "manet_global_map_update": (c_int, [_P, _P, _P, _I64, _I, _P]),
    "manet_local_map_store_select": (c_int, [_P, _P, _P, _I, _F, _P, _I64, _P]),
    "manet_correlation_output_shape": (c_int, [_I, _I, _I, _I, _I, _I, _I, _I, POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "manet_correlation_forward": (c_int, [_P, POINTER(_I64), _P, POINTER(_I64), _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "manet_correlation_backward": (c_int, [_P, POINTER(_I64), _P, POINTER(_I64), _P, _P, _P, POINTER(_I64), _P, _P,
                                           _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "manet_seghead_packed_bytes": (_SZ, []),
    "manet_seghead_workspace_bytes": (_SZ, [_I, _I, _I]),
    "manet_seghead_pack": (c_int, [POINTER(_P), _I, _I, _F, _P, _P]),
    "manet_seghead_forward": (c_int, [_P, _I, _P, POINTER(_I64), _I, _I, _I, _P, _P, _SZ, _P]),
    "manet_seghead_forward_parts": (c_int, [_P, _P, _I64, _I64, _I64, _I, _P, _P, _P, _P, _I, _I, _I, _P, _P, _SZ, _P]),
    "manet_seghead_forward_interaction": (c_int, [_P, _P, _I64, _I64, _I64, _I, _P, _P, _P, _I, _I, _I, _P, _P, _SZ, _P]),
    "manet_upsample_argmax": (c_int, [_P, _I, _I, _I, _I, _I, _P, _P, _P]),
    "manet_rough_roi": (c_int, [_P, _I, _I, _I, _I, _P, _P, _P]),
    "manet_profile_enable": (c_int, [_I]),
    "manet_profile_reset": (c_int, []),
    "manet_profile_read": (c_int, [_I, POINTER(c_float), _I, POINTER(c_int)]),
    "manet_profile_read_span": (c_int, [_I, _I, POINTER(c_float), POINTER(c_float), _I, POINTER(c_int)]),
    "manet_profile_launch_count": (ctypes.c_longlong, []),
    "manet_profile_reset_launches": (c_int, []),
    "manet_global_match_stats": (c_int, [_P, POINTER(c_int32), _P]),
    "manet_set_option": (c_int, [c_char_p, _I]),
    "manet_microbench_tmem_ld": (c_int, [_I, _I, _I, _I, POINTER(ctypes.c_longlong), _P]),
    "manet_session_create": (_P, [_I, _I, _I, _I, _I, _I]),
    "manet_session_destroy": (None, [_P]),
    "manet_session_host_buffers": (c_int, [_P] + [POINTER(_P)] * 7),
    "manet_session_step_host": (c_int, [_P, _I, _I, _I, c_uint32]),
    "manet_session_slot_buffers": (c_int, [_P, _I] + [POINTER(_P)] * 7),
    "manet_session_submit_host": (c_int, [_P, _I, _I, _I, _I, c_uint32]),
    "manet_session_wait": (c_int, [_P, _I]),
    "manet_session_upload": (c_int, [_P]),
    "manet_session_step_device": (c_int, [_P, _I, _I, _I, c_uint32]),
    "manet_session_sync": (c_int, [_P]),
    "manet_session_stream": (_P, [_P]),
}


def declared_symbols(header_path: str = HEADER_PATH):
    """Function names declared in include/manet_b200.h (used by the symbol-export test)."""
    text = open(header_path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(manet_[a-z0-9_]+)\s*\(", text)))


class ManetError(RuntimeError):
    """A libmanet_b200 call failed (the analogue of AT_ERROR("CUDA call failed"),
    correlation_cuda.cc:81-83)."""


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -m cvpr2020_manet_b200.build` "
                "(there is no CPU or PyTorch fallback for this path)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.manet_abi_version() != 1:
            raise ImportError("libmanet_b200.so ABI version mismatch; rebuild")
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().manet_last_error().decode("utf-8", "replace")
        raise ManetError(f"{what or 'libmanet_b200'} failed (code {rc}): {msg}")
