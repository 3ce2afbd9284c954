"""GPU-vs-oracle parity on REALISTIC network outputs.

tests/golden/realistic_kitti_*.npz hold the output_det / output_desc tensors of the reference's own retrained
SuperPoint model (models/sp_mbv1_b1.onnx) on its own KITTI sample images (sample_images/0000000000-3.png), generated
by tests/golden/make_realistic.py.  Unlike the N(0,1) logits used elsewhere these heatmaps are sparse and clustered
(3 % of the pixels are candidates, the greedy walk visits ~5 000 of them for 1 000 keypoints, and fewer than 2 048
survive NMS at all), which drives k_detect through its exact multi-chunk path.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    d = np.load(os.path.join(GOLD, name))
    return d["semi"], d["desc"]


def _check_decode(S, O, semi, desc, K, f16, **cfg):
    B = semi.shape[0]
    Hc, Wc = semi.shape[2:]
    fe = S.Frontend(0, B, Hc * 8, Wc * 8, K)
    if f16:
        r = fe.decode(semi, desc, max_keypoints=K, **cfg)
    else:
        r = fe.decode(semi.astype(np.float32), None if desc is None else desc.astype(np.float32), max_keypoints=K, **cfg)
    slow = int(fe.debug_counters()[0])
    fe.close()
    o = O.decode(semi.astype(np.float32), None if desc is None else desc.astype(np.float32), max_keypoints=K,
                 num_threads=8, **cfg)
    assert (r["n"] == o["n"]).all(), (r["n"], o["n"])
    for b in range(B):
        n = int(o["n"][b])
        assert (r["kpts"][b, :n] == o["kpts"][b, :n]).all(), (b, "keypoints / order")
        assert (r["scores"][b, :n].view(np.uint32) == o["scores"][b, :n].view(np.uint32)).all(), (b, "score bits")
        if desc is not None:
            assert (r["desc"][b, :n].view(np.uint32) == o["desc"][b, :n].view(np.uint32)).all(), (b, "descriptor bits")
    return o, slow


@pytest.mark.parametrize("K", [500, 1000, 2048, 4096])
@pytest.mark.parametrize("f16", [False, True])
def test_decode_kitti_1240x376(spvo, oracle, K, f16):
    semi, desc = _load("realistic_kitti_1240x376.npz")
    o, slow = _check_decode(spvo, oracle, semi[:2], desc, K, f16)
    if K >= 2048:
        assert (o["n"] < K).all(), "fewer than K survivors exist: every candidate is walked"
    o, _ = _check_decode(spvo, oracle, semi, None, K, f16)  # all four frames, detector only


@pytest.mark.parametrize("cfg", [dict(conf_thresh=0.015, dist_thresh=4, border_remove=4),
                                 dict(conf_thresh=0.001, dist_thresh=2, border_remove=0),
                                 dict(conf_thresh=0.05, dist_thresh=8, border_remove=16)])
def test_decode_kitti_784x240_parameter_sweep(spvo, oracle, cfg):
    semi, desc = _load("realistic_kitti_784x240.npz")
    _check_decode(spvo, oracle, semi, desc, 1000, False, **cfg)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_match_consecutive_frames(spvo, oracle, mode):
    """t <-> t-1 matching of real descriptors (frames 1 vs 0), both matcher algorithms."""
    S, O = spvo, oracle
    semi, desc = _load("realistic_kitti_1240x376.npz")
    o = O.decode(semi[:2].astype(np.float32), desc.astype(np.float32), max_keypoints=1000, num_threads=8)
    q, t = o["desc"][1, : int(o["n"][1])], o["desc"][0, : int(o["n"][0])]
    om, omap = O.match(q, t, mode=mode)
    assert len(om) > 300
    fe = S.Frontend(0, 2, 376, 1240, 1000)
    for alg in (S.MATCHER_TENSOR, S.MATCHER_EXACT_FP32):
        gm, gmap = fe.match(q, t, mode=mode, algorithm=alg)
        assert len(gm) == len(om) and (gm["queryIdx"] == om["queryIdx"]).all() and (gm["trainIdx"] == om["trainIdx"]).all()
        assert (gm["distance"].view(np.uint32) == om["distance"].view(np.uint32)).all() and (gmap == omap).all()
    fe.close()


def test_stereo_pipeline_on_real_frames(spvo, oracle):
    """The whole pipeline on real tensors: consecutive frames stand in for (left, right) -- the sample set has no
    right camera -- so L<->R, temporal matching, the carry and the quadruples all see real descriptor statistics."""
    from test_gpu_bench_shape import _check_frame, _oracle_frame
    S, O = spvo, oracle
    semi, desc = _load("realistic_kitti_784x240.npz")
    semi, desc = semi.astype(np.float32), desc.astype(np.float32)
    H, W, K, F = 240, 784, 1000, 2
    # frame 0 = (img0, img1), frame 1 = (img1, img0): left_1 == right_0, so temporal matches are non-trivial
    s = np.stack([semi, semi[::-1]])
    d = np.stack([desc, desc[::-1]])
    for mode in (1, 2):
        fe = S.Frontend(0, 2 * F, H, W, K)
        out = {k: v.numpy() for k, v in fe.alloc_stereo_out(F, K, device="cpu").items()}
        fe.stereo_batch(s, d, F, H, W, out, max_keypoints=K, mode=mode)
        fe.close()
        for f in range(F):
            r = _oracle_frame(O, s[max(f - 1, 0):f + 1], d[max(f - 1, 0):f + 1], f, K, mode)
            _check_frame(S, out, r, f, F, K, f"mode {mode} frame {f}")
            assert len(r["ms"]) > 200
