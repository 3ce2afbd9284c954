"""ctypes declarations for the C ABI in include/spvo_frontend.h.  Fails loudly if the CUDA library
is missing -- there is no CPU fallback in this package."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspvo_frontend.so")

SPVO_OK, SPVO_EINVAL, SPVO_ECUDA, SPVO_ENODEVICE, SPVO_ENOMEM = 0, 1, 2, 3, 4
MATCH_NN, MATCH_NN_CROSSCHECK, MATCH_KNN_RATIO = 0, 1, 2
MATCHER_AUTO, MATCHER_EXACT_FP32, MATCHER_TENSOR = 0, 1, 2
MATCH_FLAG_ROW_BAND = 1

KEYPOINT_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
     ("octave", "<i4"), ("class_id", "<i4")])
DMATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])
assert KEYPOINT_DTYPE.itemsize == 28 and DMATCH_DTYPE.itemsize == 16


class DecodeCfg(C.Structure):
    _fields_ = [("conf_thresh", C.c_float), ("dist_thresh", C.c_int32), ("border_remove", C.c_int32),
                ("max_keypoints", C.c_int32)]


class MatchCfg(C.Structure):
    _fields_ = [("mode", C.c_int32), ("ratio", C.c_float), ("algorithm", C.c_int32), ("flags", C.c_int32)]


class StereoCfg(C.Structure):
    _fields_ = [("decode", DecodeCfg), ("match", MatchCfg), ("stereo_threshold", C.c_float),
                ("min_disparity", C.c_float)]


class StereoOut(C.Structure):
    _fields_ = [("kpts", C.c_void_p), ("desc", C.c_void_p), ("n_kpts", C.c_void_p), ("matches", C.c_void_p),
                ("n_matches", C.c_void_p), ("q2t", C.c_void_p), ("stereo_keep", C.c_void_p), ("quads", C.c_void_p),
                ("n_quads", C.c_void_p)]


# every symbol include/spvo_frontend.h declares (checked by tests/test_abi.py against the header)
SYMBOLS = [
    "spvo_create", "spvo_destroy", "spvo_last_error", "spvo_abi_version", "spvo_set_stream", "spvo_sync",
    "spvo_preprocess", "spvo_preprocess_device", "spvo_decode", "spvo_decode_device", "spvo_match", "spvo_match_device", "spvo_match_masked", "spvo_match_masked_device", "spvo_match_batch_device",
    "spvo_stereo_filter_batch_device", "spvo_stereo_reset", "spvo_set_graph_mode", "spvo_stereo_batch_device", "spvo_stereo_batch",
    "spvo_decode_f16", "spvo_decode_device_f16", "spvo_stereo_batch_f16", "spvo_stereo_batch_device_f16",
    "spvo_kernel_launches", "spvo_debug_counters", "spvo_debug_div_check", "spvo_profile_enable", "spvo_profile_num_kernels",
    "spvo_profile_kernel_name", "spvo_profile_read",
]

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  This package has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    L.spvo_abi_version.restype = ci
    L.spvo_last_error.restype = C.c_char_p
    L.spvo_last_error.argtypes = [vp]
    L.spvo_create.argtypes = [C.POINTER(vp), ci, ci, ci, ci, ci]
    L.spvo_destroy.argtypes = [vp]
    L.spvo_set_stream.argtypes = [vp, vp]
    L.spvo_sync.argtypes = [vp]
    dec = [vp, vp, vp, ci, ci, ci, C.POINTER(DecodeCfg), vp, vp, vp, vp]
    L.spvo_preprocess.argtypes = [vp, vp, ci, ci, ci, ci, ci, ci, vp, vp, vp]
    L.spvo_preprocess_device.argtypes = [vp, vp, ci, ci, ci, ci, ci, ci, vp, vp, vp]
    L.spvo_decode.argtypes = dec
    L.spvo_decode_device.argtypes = dec
    L.spvo_decode_f16.argtypes = dec
    L.spvo_decode_device_f16.argtypes = dec
    mat = [vp, vp, ci, vp, ci, ci, C.POINTER(MatchCfg), vp, vp, vp]
    L.spvo_match.argtypes = mat
    L.spvo_match_device.argtypes = mat
    mmat = [vp, vp, ci, vp, ci, ci, C.POINTER(MatchCfg), vp, vp, cf, vp, vp, vp]
    L.spvo_match_masked.argtypes = mmat
    L.spvo_match_masked_device.argtypes = mmat
    L.spvo_match_masked.restype = ci
    L.spvo_match_masked_device.restype = ci
    L.spvo_match_batch_device.argtypes = [vp, vp, vp, ci, vp, vp, ci, ci, ci, C.POINTER(MatchCfg), vp, vp, vp]
    L.spvo_stereo_filter_batch_device.argtypes = [vp, vp, ci, vp, vp, ci, ci, vp, vp, cf, cf, vp]
    L.spvo_stereo_reset.argtypes = [vp]
    L.spvo_set_graph_mode.argtypes = [vp, ci]
    L.spvo_set_graph_mode.restype = ci
    ster = [vp, vp, vp, ci, ci, ci, C.POINTER(StereoCfg), C.POINTER(StereoOut)]
    L.spvo_stereo_batch_device.argtypes = ster
    L.spvo_stereo_batch.argtypes = ster
    L.spvo_stereo_batch_f16.argtypes = ster
    L.spvo_stereo_batch_device_f16.argtypes = ster
    L.spvo_kernel_launches.restype = C.c_longlong
    L.spvo_kernel_launches.argtypes = [vp]
    L.spvo_debug_counters.argtypes = [vp, vp, ci]
    L.spvo_debug_div_check.argtypes = [vp, vp, vp, C.c_longlong, C.POINTER(C.c_longlong)]
    L.spvo_debug_div_check.restype = ci
    L.spvo_profile_enable.argtypes = [vp, ci]
    L.spvo_profile_enable.restype = ci
    L.spvo_profile_num_kernels.restype = ci
    L.spvo_profile_kernel_name.restype = C.c_char_p
    L.spvo_profile_kernel_name.argtypes = [ci]
    L.spvo_profile_read.argtypes = [vp, vp, vp, ci]
    L.spvo_profile_read.restype = ci
    for name in ("spvo_create", "spvo_destroy", "spvo_set_stream", "spvo_sync", "spvo_decode", "spvo_decode_device",
                 "spvo_match", "spvo_match_device", "spvo_match_masked", "spvo_match_masked_device", "spvo_match_batch_device", "spvo_stereo_filter_batch_device",
                 "spvo_debug_counters", "spvo_stereo_reset", "spvo_stereo_batch_device", "spvo_stereo_batch"):
        getattr(L, name).restype = ci
    _lib = L
    return L
