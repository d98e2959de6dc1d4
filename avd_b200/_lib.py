"""ctypes binding of csrc/libavd.so (C ABI: include/avd.h).  No fallback: if the CUDA extension is
missing or no B200 is visible, every entry raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AVD_LIB") or os.path.join(_HERE, "csrc", "libavd.so")   # AVD_LIB: A/B kernel variants

THUMB = 320
REC_DTYPE = np.dtype([("lap_sum", "<i8"), ("lap_sumsq", "<i8"), ("flow_mean", "<f4"), ("flow_var", "<f4"),
                      ("hash", "<u4", (32,)), ("ham_prev", "<i4"), ("flags", "<u4")])
assert REC_DTYPE.itemsize == 160
REC_HAS_PREV, REC_HAS_FLOW = 1, 2

SYMBOLS = ["avd_version", "avd_device_count", "avd_create", "avd_destroy", "avd_last_error", "avd_sync",
           "avd_host_alloc", "avd_host_free", "avd_frames_features", "avd_flow_stats", "avd_flow_fields",
           "avd_gray", "avd_analyze_host", "avd_launch_count", "avd_set_k1_mode", "avd_profile_enable",
           "avd_profile_read", "avd_shape_plan", "avd_score_clips", "avd_audio_windows", "avd_ws_generation",
           "avd_host_alloc_count", "avd_yuv420_to_bgr", "avd_analyze_host_yuv420", "avd_check_guards"]


class AvdError(RuntimeError):
    pass


_lib = None


def load():
    """Load libavd.so and declare prototypes.  Raises AvdError if the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AvdError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(nvcc, sm_100a).  There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u8p, sz = C.c_void_p, C.c_void_p, C.c_size_t
    L.avd_version.restype = C.c_char_p
    L.avd_device_count.restype = C.c_int
    L.avd_create.argtypes = [C.c_int, C.POINTER(vp)]; L.avd_create.restype = C.c_int
    L.avd_destroy.argtypes = [vp]; L.avd_destroy.restype = None
    L.avd_last_error.argtypes = [vp]; L.avd_last_error.restype = C.c_char_p
    L.avd_sync.argtypes = [vp]; L.avd_sync.restype = C.c_int
    L.avd_host_alloc.argtypes = [vp, sz]; L.avd_host_alloc.restype = vp
    L.avd_host_free.argtypes = [vp, vp]; L.avd_host_free.restype = None
    L.avd_frames_features.argtypes = [vp, u8p, C.c_int, C.c_int, C.c_int, sz, sz, vp, vp, u8p, u8p, vp]
    L.avd_frames_features.restype = C.c_int
    L.avd_flow_stats.argtypes = [vp, u8p, C.c_int, u8p, vp, vp]; L.avd_flow_stats.restype = C.c_int
    L.avd_flow_fields.argtypes = [vp, u8p, u8p, C.c_int, vp, vp]; L.avd_flow_fields.restype = C.c_int
    L.avd_gray.argtypes = [vp, u8p, C.c_int, C.c_int, sz, u8p, vp]; L.avd_gray.restype = C.c_int
    L.avd_analyze_host.argtypes = [vp, C.POINTER(vp), C.c_int, C.c_int, C.c_int, sz, vp]; L.avd_analyze_host.restype = C.c_int
    L.avd_launch_count.argtypes = [vp]; L.avd_launch_count.restype = C.c_uint64
    L.avd_set_k1_mode.argtypes = [vp, C.c_int]; L.avd_set_k1_mode.restype = C.c_int
    L.avd_profile_enable.argtypes = [vp, C.c_int]; L.avd_profile_enable.restype = C.c_int
    L.avd_profile_read.argtypes = [vp, C.c_char_p, sz]; L.avd_profile_read.restype = C.c_int
    L.avd_score_clips.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, vp, vp, vp, vp, C.c_int, vp]
    L.avd_score_clips.restype = C.c_int
    L.avd_audio_windows.argtypes = [vp, vp, C.c_longlong, C.c_int, vp, vp]; L.avd_audio_windows.restype = C.c_int
    L.avd_ws_generation.argtypes = [vp]; L.avd_ws_generation.restype = C.c_uint64
    L.avd_host_alloc_count.argtypes = [vp]; L.avd_host_alloc_count.restype = sz
    L.avd_check_guards.argtypes = [vp]; L.avd_check_guards.restype = C.c_longlong
    L.avd_yuv420_to_bgr.argtypes = [vp, u8p, sz, C.c_int, C.c_int, C.c_int, C.c_int, u8p, sz, sz, vp]; L.avd_yuv420_to_bgr.restype = C.c_int
    L.avd_analyze_host_yuv420.argtypes = [vp, C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, vp]; L.avd_analyze_host_yuv420.restype = C.c_int
    _lib = L
    return L
