"""ctypes wrapper around oracle/libspvo_oracle.so -- the CPU ORACLE (test infrastructure only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package never does.  See spvo_oracle.cpp for the reference file:line
citations of every function.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libspvo_oracle.so")

KEYPOINT_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
     ("octave", "<i4"), ("class_id", "<i4")])
DMATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])
assert KEYPOINT_DTYPE.itemsize == 28 and DMATCH_DTYPE.itemsize == 16

MODE_NN, MODE_NN_CROSSCHECK, MODE_KNN_RATIO = 0, 1, 2


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (g++).  Building the checker is not using it."""
    src = os.path.join(_HERE, "spvo_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libspvo_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        fp, ip, vp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_void_p
        L.spvo_oracle_exp.restype = C.c_float
        L.spvo_oracle_exp.argtypes = [C.c_float]
        L.spvo_oracle_l2dist.restype = C.c_float
        L.spvo_oracle_l2dist.argtypes = [vp, vp, C.c_int]
        L.spvo_oracle_heatmap.restype = C.c_int
        L.spvo_oracle_heatmap.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int]
        L.spvo_oracle_heatmap_variant.restype = C.c_int
        L.spvo_oracle_heatmap_variant.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int]
        L.spvo_oracle_detect.restype = C.c_int
        L.spvo_oracle_detect.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                                         vp, vp, vp]
        L.spvo_oracle_decode.restype = C.c_int
        L.spvo_oracle_decode.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,
                                         C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, C.c_int]
        L.spvo_oracle_match.restype = C.c_int
        L.spvo_oracle_match.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_float, vp, vp, vp,
                                        C.c_int]
        L.spvo_oracle_match_masked.restype = C.c_int
        L.spvo_oracle_match_masked.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_float, vp, vp, C.c_float,
                                               vp, vp, vp, C.c_int]
        L.spvo_oracle_stereo_filter.restype = C.c_int
        L.spvo_oracle_stereo_filter.argtypes = [vp, vp, vp, C.c_int, C.c_float, C.c_float, vp]
        L.spvo_oracle_consistency.restype = C.c_int
        L.spvo_oracle_consistency.argtypes = [vp, C.c_int, vp, vp, vp, vp]
        L.spvo_oracle_crop_geometry.restype = C.c_int
        L.spvo_oracle_crop_geometry.argtypes = [C.c_int] * 4 + [ip] * 4
        L.spvo_oracle_preprocess.restype = C.c_int
        L.spvo_oracle_preprocess.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def exp(x) -> np.ndarray:
    x = _f32(np.atleast_1d(x))
    L = lib()
    return np.array([L.spvo_oracle_exp(float(v)) for v in x.ravel()], dtype=np.float32).reshape(x.shape)


def l2dist(a, b) -> float:
    a, b = _f32(a), _f32(b)
    return float(lib().spvo_oracle_l2dist(_p(a), _p(b), a.size))


def heatmap(semi, num_threads: int = 1) -> np.ndarray:
    """semi [B,65,Hc,Wc] -> heat [B,8Hc,8Wc]  (rows D1-D2)."""
    semi = _f32(semi)
    B, Cc, Hc, Wc = semi.shape
    assert Cc == 65
    heat = np.empty((B, Hc * 8, Wc * 8), np.float32)
    rc = lib().spvo_oracle_heatmap(_p(semi), B, Hc * 8, Wc * 8, _p(heat), num_threads)
    assert rc == 0
    return heat


VAR_LIBM_EXP, VAR_CEPHES_NOFMA, VAR_PAIRWISE_SUM = 1, 2, 4


def heatmap_variant(semi, variant: int, num_threads: int = 1) -> np.ndarray:
    """Sensitivity study only: the heatmap with another plausible exp / channel-sum build (variant 0 = the spec)."""
    semi = _f32(semi)
    B, Cc, Hc, Wc = semi.shape
    heat = np.empty((B, Hc * 8, Wc * 8), np.float32)
    assert lib().spvo_oracle_heatmap_variant(_p(semi), B, Hc * 8, Wc * 8, _p(heat), num_threads, int(variant)) == 0
    return heat


def detect(heat, conf_thresh=0.015, dist_thresh=4, border_remove=4, max_keypoints=1000, faithful_sort=False):
    """Rows D3-D5 (threshold, sort, greedy NMS, border, top-K) on a given heatmap [B,H,W]."""
    heat = _f32(heat)
    B, H, W = heat.shape
    K = int(max_keypoints)
    kp = np.zeros((B, K), KEYPOINT_DTYPE)
    sc = np.zeros((B, K), np.float32)
    n = np.zeros(B, np.int32)
    assert lib().spvo_oracle_detect(_p(heat), B, H, W, float(conf_thresh), int(dist_thresh), int(border_remove), K,
                                    int(bool(faithful_sort)), _p(kp), _p(sc), _p(n)) == 0
    return dict(kpts=kp, scores=sc, n=n)


def decode(semi, desc, conf_thresh=0.015, dist_thresh=4, border_remove=4, max_keypoints=1000,
           faithful_sort=False, num_threads=1, want_heat=False):
    """Full decode of a batch.  Returns dict(kpts [B,K] structured, desc [B,K,256], n [B], scores [B,K],
    walked [B], ncand [B], heat (optional))."""
    semi = _f32(semi)
    B, Cc, Hc, Wc = semi.shape
    assert Cc == 65
    H, W, K = Hc * 8, Wc * 8, int(max_keypoints)
    if desc is not None:
        desc = _f32(desc)
        assert desc.shape == (B, 256, Hc, Wc)
    kp = np.zeros((B, max(K, 1)), KEYPOINT_DTYPE)[:, :K].copy() if K == 0 else np.zeros((B, K), KEYPOINT_DTYPE)
    dout = np.zeros((B, K, 256), np.float32) if desc is not None else None
    n = np.zeros(B, np.int32)
    scores = np.zeros((B, K), np.float32)
    walked = np.zeros(B, np.int32)
    ncand = np.zeros(B, np.int32)
    heat = np.empty((B, H, W), np.float32) if want_heat else None
    rc = lib().spvo_oracle_decode(_p(semi), _p(desc), B, H, W, float(conf_thresh), int(dist_thresh),
                                  int(border_remove), K, int(bool(faithful_sort)), _p(kp), _p(dout), _p(n),
                                  _p(scores), _p(heat), _p(walked), _p(ncand), int(num_threads))
    if rc != 0:
        raise ValueError("spvo_oracle_decode: invalid arguments")
    return dict(kpts=kp, desc=dout, n=n, scores=scores, walked=walked, ncand=ncand, heat=heat)


def match(q, t, mode=MODE_NN_CROSSCHECK, ratio=0.8, num_threads=1, qy=None, ty=None, band=-1.0):
    """Returns (matches structured [n], q2t [N]).  qy / ty / band: optional row-band mask |qy[i] - ty[j]| <= band."""
    q = _f32(q).reshape(-1, q.shape[-1] if q.ndim > 1 else 256)
    t = _f32(t).reshape(-1, t.shape[-1] if t.ndim > 1 else 256)
    N, D = q.shape
    M = t.shape[0]
    out = np.zeros(max(N, 1), DMATCH_DTYPE)
    q2t = np.full(max(N, 1), -1, np.int32)
    n = C.c_int(0)
    if qy is not None:
        qy, ty = _f32(qy).reshape(-1), _f32(ty).reshape(-1)
        assert qy.size == N and ty.size == M
    rc = lib().spvo_oracle_match_masked(_p(q), N, _p(t), M, D, int(mode), float(ratio), _p(qy), _p(ty), float(band),
                                        _p(out), C.byref(n), _p(q2t), int(num_threads))
    if rc != 0:
        raise ValueError("spvo_oracle_match: invalid arguments")
    return out[: n.value].copy(), q2t[:N].copy()


def stereo_filter(kl, kr, matches, stereo_threshold=2.0, min_disparity=0.25) -> np.ndarray:
    kl = np.ascontiguousarray(kl, KEYPOINT_DTYPE)
    kr = np.ascontiguousarray(kr, KEYPOINT_DTYPE)
    m = np.ascontiguousarray(matches, DMATCH_DTYPE)
    keep = np.zeros(max(len(m), 1), np.uint8)
    lib().spvo_oracle_stereo_filter(_p(kl), _p(kr), _p(m), len(m), float(stereo_threshold),
                                    float(min_disparity), _p(keep))
    return keep[: len(m)].astype(bool)


def consistency(stereo_matches, map_t, keep, map_prev) -> np.ndarray:
    """Quadruples (currL, currR, prevL, prevR) of BASE:156-207; returns int32 [n, 4]."""
    m = np.ascontiguousarray(stereo_matches, DMATCH_DTYPE)
    mt = np.ascontiguousarray(map_t, np.int32)
    mp = np.ascontiguousarray(map_prev, np.int32)
    kp = np.ascontiguousarray(keep, np.uint8)
    out = np.zeros((max(len(m), 1), 4), np.int32)
    n = lib().spvo_oracle_consistency(_p(m), len(m), _p(mt), _p(kp), _p(mp), _p(out))
    return out[:n].copy()


def crop_geometry(rows, cols, H, W):
    """(crop_rows, crop_cols, row_offset, col_offset) of BASE:71-113."""
    v = [C.c_int(0) for _ in range(4)]
    if lib().spvo_oracle_crop_geometry(rows, cols, H, W, *[C.byref(x) for x in v]):
        raise ValueError("spvo_oracle_crop_geometry: invalid arguments")
    return tuple(x.value for x in v)


def preprocess(img, H, W, P=None):
    """img [rows, cols] uint8 -> (input [H,W] float32, resized [H,W] uint8, patched P [3,4] or None)."""
    img = np.ascontiguousarray(img, np.uint8)
    rows, cols = img.shape
    inp = np.empty((H, W), np.float32)
    rs = np.empty((H, W), np.uint8)
    Pm = None if P is None else np.ascontiguousarray(P, np.float32).reshape(3, 4).copy()
    if lib().spvo_oracle_preprocess(_p(img), rows, cols, cols, H, W, _p(inp), _p(rs), _p(Pm)):
        raise ValueError("spvo_oracle_preprocess: invalid arguments")
    return inp, rs, Pm
