"""Host-side mirror of the reference's front-end interface, on top of the C ABI.

`Frontend` is a thin handle wrapper (one per GPU / stream).  `SuperPointFeatureFrontEnd` mirrors the
reference class of the same name (include/odml_visual_odometry/feature_detection.hpp:253-391 and
:96-178): same member names (`keypoints_dq`, `descriptors_dq`, `cv_DMatches_list`,
`maps_of_indices`), same call sequence (`postprocessDetectionAndDescription()` after the network
wrote `output_det_data_` / `output_desc_data_`, then `matchDescriptors(match_type)`), so the parity
tests read like the reference's call sites (visual_odometry_node.cpp:175-199).

All compute goes through libspvo_frontend.so (hand-written sm_100a CUDA).  No CPU fallback.
"""
from __future__ import annotations

import collections
import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from ._lib import (DMATCH_DTYPE, KEYPOINT_DTYPE, MATCH_KNN_RATIO, MATCH_NN, MATCH_NN_CROSSCHECK, MATCHER_AUTO,
                   DecodeCfg, MatchCfg)

# reference enums (feature_detection.hpp:66-90)
PREV_LEFT, PREV_RIGHT, CURR_LEFT, CURR_RIGHT = -4, -3, -2, -1
CURR_LEFT_CURR_RIGHT, CURR_LEFT_PREV_LEFT, PREV_LEFT_PREV_RIGHT, MATCH_TYPE_NUM = 0, 1, 2, 3
match_type_to_positions = ((CURR_LEFT, CURR_RIGHT), (CURR_LEFT, PREV_LEFT), (PREV_LEFT, PREV_RIGHT))


class SpvoError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"spvo error {code}: {msg}")
        self.code = code


def _ptr(x) -> Optional[int]:
    """Device pointer of a torch tensor / raw int, or host pointer of a numpy array."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return x.data_ptr()  # torch.Tensor


class Frontend:
    """One library handle: owns workspaces for up to max_batch images of max_height x max_width."""

    def __init__(self, device: int = 0, max_batch: int = 2, max_height: int = 376, max_width: int = 1240,
                 max_keypoints: int = 1000):
        self._L = _lib.load()
        self._h = C.c_void_p()
        rc = self._L.spvo_create(C.byref(self._h), device, max_batch, max_height, max_width, max_keypoints)
        if rc != 0:
            raise SpvoError(rc, self._L.spvo_last_error(None).decode())
        self.device, self.max_batch, self.max_keypoints = device, max_batch, max_keypoints

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.spvo_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def _check(self, rc: int):
        if rc != 0:
            raise SpvoError(rc, self._L.spvo_last_error(self._h).decode())

    def set_stream(self, cuda_stream: int = 0):
        self._check(self._L.spvo_set_stream(self._h, cuda_stream or None))

    def sync(self):
        self._check(self._L.spvo_sync(self._h))

    @property
    def kernel_launches(self) -> int:
        return int(self._L.spvo_kernel_launches(self._h))

    def debug_counters(self) -> np.ndarray:
        out = np.zeros(8, np.int64)
        self._check(self._L.spvo_debug_counters(self._h, out.ctypes.data, 8))
        return out

    def div_check(self, a_bits, b_bits) -> int:
        """Mismatches between k_softmax_heat's shared-reciprocal division and the IEEE division on the given operand
        bit patterns (CUDA int32/uint32 tensors of equal length).  Test hook; must return 0."""
        n = C.c_longlong(0)
        self._check(self._L.spvo_debug_div_check(self._h, _ptr(a_bits), _ptr(b_bits), a_bits.numel(), C.byref(n)))
        return int(n.value)

    def profile_enable(self, on: bool = True):
        self._check(self._L.spvo_profile_enable(self._h, int(on)))

    def profile_read(self) -> dict:
        """{kernel name: (total ms, launches)} since the last read (synchronises the stream)."""
        n = self._L.spvo_profile_num_kernels()
        ms, cnt = np.zeros(n, np.float64), np.zeros(n, np.int64)
        self._check(self._L.spvo_profile_read(self._h, ms.ctypes.data, cnt.ctypes.data, n))
        return {self._L.spvo_profile_kernel_name(i).decode(): (float(ms[i]), int(cnt[i]))
                for i in range(n) if cnt[i] > 0}

    # ---- preprocess ---------------------------------------------------------------------------
    def preprocess(self, imgs: np.ndarray, H: int, W: int, proj: Optional[np.ndarray] = None):
        """Host-buffer preprocess (spvo_preprocess).  imgs [B,rows,cols] uint8 ->
        (input [B,H,W] float32, resized [B,H,W] uint8, patched proj [B,3,4] or None)."""
        imgs = np.ascontiguousarray(imgs, np.uint8)
        if imgs.ndim == 2:
            imgs = imgs[None]
        B, rows, cols = imgs.shape
        inp = np.empty((B, H, W), np.float32)
        rs = np.empty((B, H, W), np.uint8)
        P = None if proj is None else np.ascontiguousarray(proj, np.float32).reshape(B, 12).copy()
        self._check(self._L.spvo_preprocess(self._h, _ptr(imgs), B, rows, cols, cols, H, W, _ptr(inp), _ptr(rs),
                                            _ptr(P)))
        return inp, rs, (None if P is None else P.reshape(B, 3, 4))

    def preprocess_device(self, imgs, B, rows, cols, stride, H, W, input_out, resized_out=None, proj=None):
        """Device-pointer preprocess (spvo_preprocess_device): torch CUDA tensors or raw pointers, async;
        `proj` (numpy float32 [B,12], host) is patched in place."""
        self._check(self._L.spvo_preprocess_device(self._h, _ptr(imgs), B, rows, cols, stride, H, W,
                                                   _ptr(input_out), _ptr(resized_out), _ptr(proj)))

    # ---- decode -------------------------------------------------------------------------------
    def decode(self, semi: np.ndarray, desc: Optional[np.ndarray], conf_thresh=0.015, dist_thresh=4,
               border_remove=4, max_keypoints=1000, want_scores=True):
        """Host-buffer decode (spvo_decode / spvo_decode_f16).  semi [B,65,Hc,Wc], desc [B,256,Hc,Wc] numpy fp32, or
        both float16 (an fp16 engine's outputs; results equal those of the widened fp32 tensors bit for bit)."""
        f16 = np.asarray(semi).dtype == np.float16
        dt = np.float16 if f16 else np.float32
        semi = np.ascontiguousarray(semi, dt)
        B, Cc, Hc, Wc = semi.shape
        if Cc != 65:
            raise ValueError("semi must be [B,65,Hc,Wc]")
        K = int(max_keypoints)
        if desc is not None:
            desc = np.ascontiguousarray(desc, dt)
            if desc.shape != (B, 256, Hc, Wc):
                raise ValueError("desc must be [B,256,Hc,Wc]")
        kp = np.zeros((B, K), KEYPOINT_DTYPE)
        dout = np.zeros((B, K, 256), np.float32) if desc is not None else None
        n = np.zeros(B, np.int32)
        sc = np.zeros((B, K), np.float32) if want_scores else None
        cfg = DecodeCfg(conf_thresh, dist_thresh, border_remove, K)
        fn = self._L.spvo_decode_f16 if f16 else self._L.spvo_decode
        self._check(fn(self._h, _ptr(semi), _ptr(desc), B, Hc * 8, Wc * 8, C.byref(cfg), _ptr(kp), _ptr(dout), _ptr(n),
                       _ptr(sc)))
        return dict(kpts=kp, desc=dout, n=n, scores=sc)

    def decode_device(self, semi, desc, B, H, W, kpts_out, desc_out, n_out, scores_out=None, conf_thresh=0.015,
                      dist_thresh=4, border_remove=4, max_keypoints=1000, f16=False):
        """Device-pointer decode (spvo_decode_device[_f16]): torch CUDA tensors or raw pointers; async."""
        cfg = DecodeCfg(conf_thresh, dist_thresh, border_remove, int(max_keypoints))
        fn = self._L.spvo_decode_device_f16 if f16 else self._L.spvo_decode_device
        self._check(fn(self._h, _ptr(semi), _ptr(desc), B, H, W, C.byref(cfg), _ptr(kpts_out), _ptr(desc_out),
                       _ptr(n_out), _ptr(scores_out)))

    # ---- match --------------------------------------------------------------------------------
    def match(self, q: np.ndarray, t: np.ndarray, mode=MATCH_NN_CROSSCHECK, ratio=0.8, algorithm=MATCHER_AUTO,
              q_kpts=None, t_kpts=None, band=None):
        """Host-buffer match (spvo_match; spvo_match_masked when q_kpts / t_kpts / band give a row-band mask).
        Returns (DMatch structured array, q2t map)."""
        q = np.ascontiguousarray(q, np.float32).reshape(-1, 256)
        t = np.ascontiguousarray(t, np.float32).reshape(-1, 256)
        N, M = q.shape[0], t.shape[0]
        out = np.zeros(max(N, 1), DMATCH_DTYPE)
        q2t = np.full(max(N, 1), -1, np.int32)
        n = C.c_int(0)
        cfg = MatchCfg(mode, ratio, algorithm, 0)
        if band is None:
            self._check(self._L.spvo_match(self._h, _ptr(q), N, _ptr(t), M, 256, C.byref(cfg), _ptr(out),
                                           C.addressof(n), _ptr(q2t)))
        else:
            qk = np.ascontiguousarray(q_kpts, KEYPOINT_DTYPE)
            tk = np.ascontiguousarray(t_kpts, KEYPOINT_DTYPE)
            assert len(qk) == N and len(tk) == M
            self._check(self._L.spvo_match_masked(self._h, _ptr(q), N, _ptr(t), M, 256, C.byref(cfg), _ptr(qk), _ptr(tk),
                                                  float(band), _ptr(out), C.addressof(n), _ptr(q2t)))
        return out[: n.value].copy(), q2t[:N].copy()

    def match_device(self, q, N, t, M, out, n_matches, q2t=None, mode=MATCH_NN_CROSSCHECK, ratio=0.8,
                     algorithm=MATCHER_AUTO):
        cfg = MatchCfg(mode, ratio, algorithm, 0)
        self._check(self._L.spvo_match_device(self._h, _ptr(q), N, _ptr(t), M, 256, C.byref(cfg), _ptr(out),
                                              _ptr(n_matches), _ptr(q2t)))

    def match_batch_device(self, desc_base, n_rows, slot_stride_rows, q_slot, t_slot, P, max_rows, out, n_matches,
                           q2t=None, mode=MATCH_NN_CROSSCHECK, ratio=0.8, algorithm=MATCHER_AUTO):
        cfg = MatchCfg(mode, ratio, algorithm, 0)
        self._check(self._L.spvo_match_batch_device(self._h, _ptr(desc_base), _ptr(n_rows), slot_stride_rows,
                                                    _ptr(q_slot), _ptr(t_slot), P, max_rows, 256, C.byref(cfg),
                                                    _ptr(out), _ptr(n_matches), _ptr(q2t)))

    def stereo_filter_batch_device(self, kpts_base, slot_stride_rows, q_slot, t_slot, P, max_rows, matches,
                                   n_matches, keep, stereo_threshold=2.0, min_disparity=0.25):
        self._check(self._L.spvo_stereo_filter_batch_device(self._h, _ptr(kpts_base), slot_stride_rows, _ptr(q_slot),
                                                            _ptr(t_slot), P, max_rows, _ptr(matches), _ptr(n_matches),
                                                            stereo_threshold, min_disparity, _ptr(keep)))


    # ---- stereo stream ----------------------------------------------------------------------
    def stereo_reset(self):
        self._check(self._L.spvo_stereo_reset(self._h))

    def set_graph_mode(self, on: bool = True):
        """CUDA-graph replay of stereo_batch_device calls with an unchanged signature (spvo_set_graph_mode)."""
        self._check(self._L.spvo_set_graph_mode(self._h, int(on)))

    @staticmethod
    def _stereo_cfg(conf_thresh, dist_thresh, border_remove, max_keypoints, mode, ratio, algorithm,
                    stereo_threshold, min_disparity, row_band=False):
        return _lib.StereoCfg(DecodeCfg(conf_thresh, dist_thresh, border_remove, int(max_keypoints)),
                              MatchCfg(mode, ratio, algorithm, _lib.MATCH_FLAG_ROW_BAND if row_band else 0),
                              stereo_threshold, min_disparity)

    def stereo_batch_device(self, semi, desc, F, H, W, out: dict, conf_thresh=0.015, dist_thresh=4, border_remove=4,
                            max_keypoints=1000, mode=MATCH_NN_CROSSCHECK, ratio=0.8, algorithm=MATCHER_AUTO,
                            stereo_threshold=2.0, min_disparity=0.25, f16=False, row_band=False):
        """spvo_stereo_batch_device[_f16].  `out` maps the spvo_stereo_out field names to CUDA tensors; f16=True:
        semi / desc are float16 tensors (an fp16 engine's output bindings); row_band=True: the L<->R matching runs
        under the row-band mask |y_l - y_r| <= stereo_threshold (SPVO_MATCH_FLAG_ROW_BAND)."""
        cfg = self._stereo_cfg(conf_thresh, dist_thresh, border_remove, max_keypoints, mode, ratio, algorithm,
                               stereo_threshold, min_disparity, row_band)
        so = _lib.StereoOut(*[_ptr(out.get(k)) for k in
                              ("kpts", "desc", "n_kpts", "matches", "n_matches", "q2t", "stereo_keep", "quads", "n_quads")])
        fn = self._L.spvo_stereo_batch_device_f16 if f16 else self._L.spvo_stereo_batch_device
        self._check(fn(self._h, _ptr(semi), _ptr(desc), F, H, W, C.byref(cfg), C.byref(so)))

    def stereo_batch(self, semi, desc, F, H, W, out: dict, conf_thresh=0.015, dist_thresh=4, border_remove=4,
                     max_keypoints=1000, mode=MATCH_NN_CROSSCHECK, ratio=0.8, algorithm=MATCHER_AUTO,
                     stereo_threshold=2.0, min_disparity=0.25, f16=False, row_band=False):
        """spvo_stereo_batch[_f16] (host pointers: numpy arrays or pinned CPU torch tensors); synchronous."""
        cfg = self._stereo_cfg(conf_thresh, dist_thresh, border_remove, max_keypoints, mode, ratio, algorithm,
                               stereo_threshold, min_disparity, row_band)
        so = _lib.StereoOut(*[_ptr(out.get(k)) for k in
                              ("kpts", "desc", "n_kpts", "matches", "n_matches", "q2t", "stereo_keep", "quads", "n_quads")])
        fn = self._L.spvo_stereo_batch_f16 if f16 else self._L.spvo_stereo_batch
        self._check(fn(self._h, _ptr(semi), _ptr(desc), F, H, W, C.byref(cfg), C.byref(so)))

    @staticmethod
    def alloc_stereo_out(F, K, device="cuda", pinned=False, with_desc=True):
        """Output buffers of spvo_stereo_out as torch tensors (device or pinned host)."""
        import torch
        kw = dict(device=device) if device != "cpu" else dict(pin_memory=pinned)
        out = dict(
            kpts=torch.zeros(2 * F, K, 7, dtype=torch.float32, **kw),
            n_kpts=torch.zeros(2 * F, dtype=torch.int32, **kw),
            matches=torch.zeros(2 * F, K, 4, dtype=torch.int32, **kw),
            n_matches=torch.zeros(2 * F, dtype=torch.int32, **kw),
            q2t=torch.zeros(2 * F, K, dtype=torch.int32, **kw),
            stereo_keep=torch.zeros(F, K, dtype=torch.uint8, **kw),
            quads=torch.zeros(F, K, 4, dtype=torch.int32, **kw),
            n_quads=torch.zeros(F, dtype=torch.int32, **kw),
        )
        if with_desc:
            out["desc"] = torch.zeros(2 * F, K, 256, dtype=torch.float32, **kw)
        return out


class SuperPointFeatureFrontEnd:
    """Python mirror of the reference class (feature_detection.hpp:253-391) for the decode + match path.

    The TensorRT runner is out of scope: the caller writes the network outputs into
    `output_det_data_` [B,65,H/8,W/8] and `output_desc_data_` [B,256,H/8,W/8] (hpp:383-384) and calls
    `postprocessDetectionAndDescription()`, exactly where the reference does (NN:471/475/484).
    """

    knn_threshold_ = 0.8  # hpp:137
    max_keypoints_ = 1000  # hpp:368 (runtime here)

    def __init__(self, selector_type: str = "NN", cross_check: bool = True, model_batch_size: int = 2,
                 input_height: int = 120, input_width: int = 392, conf_thresh: float = 0.015, dist_thresh: int = 4,
                 border_remove: int = 4, stereo_threshold: float = 2.0, min_disparity: float = 1.0,
                 max_keypoints: int = 1000, device: int = 0, matcher_algorithm: int = MATCHER_AUTO):
        if input_height % 8 or input_width % 8:  # hpp:296
            raise ValueError("input_height and input_width must be multiples of 8")
        if selector_type not in ("NN", "KNN"):
            raise ValueError("selector_type must be NN or KNN")  # hpp:60-64
        self.selector_type_, self.cross_check_ = selector_type, bool(cross_check)
        self.model_batch_size_ = model_batch_size
        self.input_height_, self.input_width_ = input_height, input_width
        self.output_height_, self.output_width_ = input_height // 8, input_width // 8
        self.conf_thresh_, self.dist_thresh_, self.border_remove_ = conf_thresh, dist_thresh, border_remove
        self.stereo_threshold_, self.min_disparity_ = stereo_threshold, min_disparity
        self.max_keypoints_ = max_keypoints
        self.matcher_algorithm_ = matcher_algorithm
        # initMatcher() (BASE:10-33): BFMatcher(NORM_L2, cross_check && selector != KNN)
        if selector_type == "KNN":
            self._mode = MATCH_KNN_RATIO
        else:
            self._mode = MATCH_NN_CROSSCHECK if cross_check else MATCH_NN
        # initPointers() (hpp:309-318): host I/O buffers the network writes into
        B, Hc, Wc = model_batch_size, self.output_height_, self.output_width_
        self.input_data_ = np.zeros((B, input_height, input_width), np.float32)  # hpp:382
        self.images_dq = collections.deque()                                       # hpp:123
        self.output_det_data_ = np.zeros((B, 65, Hc, Wc), np.float32)
        self.output_desc_data_ = np.zeros((B, 256, Hc, Wc), np.float32)
        self._fe = Frontend(device, B, input_height, input_width, max_keypoints)
        self.keypoints_dq = collections.deque()    # hpp:124
        self.descriptors_dq = collections.deque()  # hpp:128
        self.cv_DMatches_list = [np.zeros(0, DMATCH_DTYPE) for _ in range(MATCH_TYPE_NUM)]  # hpp:129
        self.maps_of_indices = [np.zeros(0, np.int32) for _ in range(MATCH_TYPE_NUM)]       # hpp:161

    def clearLagecyData(self):  # BASE:35-66 (sic)
        self.images_dq.clear()       # BASE:36
        self.keypoints_dq.clear()
        self.descriptors_dq.clear()
        self.cv_DMatches_list = [np.zeros(0, DMATCH_DTYPE) for _ in range(MATCH_TYPE_NUM)]
        self.maps_of_indices = [np.zeros(0, np.int32) for _ in range(MATCH_TYPE_NUM)]

    def preprocessImage(self, img: np.ndarray, projection_matrix: np.ndarray, curr_batch: int):  # NN:139-161
        """img: uint8 [rows, cols]; projection_matrix: float32 [3,4], patched in place (BASE:68-121)."""
        assert curr_batch < self.model_batch_size_
        inp, rs, P = self._fe.preprocess(img, self.input_height_, self.input_width_, projection_matrix)
        projection_matrix[...] = P[0]
        self.images_dq.append(rs[0])           # NN:153
        self.input_data_[curr_batch] = inp[0]  # NN:159-160

    def postprocessDetectionAndDescription(self):  # NN:264-364
        r = self._fe.decode(self.output_det_data_, self.output_desc_data_, self.conf_thresh_, self.dist_thresh_,
                            self.border_remove_, self.max_keypoints_, want_scores=False)
        for b in range(self.model_batch_size_):
            n = int(r["n"][b])
            self.keypoints_dq.append(r["kpts"][b, :n].copy())    # NN:261
            self.descriptors_dq.append(r["desc"][b, :n].copy())  # NN:362
        while len(self.keypoints_dq) > 4:                        # NN:494-498
            self.keypoints_dq.popleft()
            self.descriptors_dq.popleft()
        while len(self.images_dq) > 4:                           # NN:494-498 trims all three deques together
            self.images_dq.popleft()

    def matchDescriptors(self, match_type: int):  # BASE:434-500
        p0, p1 = match_type_to_positions[match_type]
        d0, d1 = self.descriptors_dq[p0], self.descriptors_dq[p1]
        k0 = self.keypoints_dq[p0]
        matches, q2t = self._fe.match(d0, d1, self._mode, self.knn_threshold_, self.matcher_algorithm_)
        self.cv_DMatches_list[match_type] = matches
        if match_type == CURR_LEFT_CURR_RIGHT:  # BASE:475-481
            self.maps_of_indices[PREV_LEFT_PREV_RIGHT] = self.maps_of_indices[CURR_LEFT_CURR_RIGHT]
        assert len(q2t) == len(k0)
        self.maps_of_indices[match_type] = q2t  # BASE:483-491
