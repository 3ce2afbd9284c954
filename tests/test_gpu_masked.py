"""Row-band MASKED matching (BASELINE north star: "left<->right matching under a stereo row-band constraint"), an
explicit opt-in next to the reference's unmasked matching + post-filter (feature_detection_base.cpp:169-172).
Oracle = tests/test_oracle_match.py::test_masked_matching_equals_masked_cv2 pins it to masked cv2.BFMatcher calls."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _kp(S, y):
    k = np.zeros(len(y), S.KEYPOINT_DTYPE)
    k["y"] = y
    k["x"] = np.arange(len(y)) % 97
    return k


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("shape", [(1000, 1000, 2.0), (300, 700, 0.5), (1500, 900, 6.0), (64, 40, 1.0), (2048, 2048, 2.0)])
def test_masked_match_both_algorithms(spvo, oracle, mode, shape):
    from conftest import unit_rows
    S, O = spvo, oracle
    N, M, band = shape
    rng = np.random.default_rng(N + M)
    q, t = unit_rows(N, seed=N), unit_rows(M, seed=M + 7)
    n_true = min(N, M) * 2 // 3
    t[:n_true] = q[:n_true] + 0.05 * rng.standard_normal((n_true, 256)).astype(np.float32)
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    qy = rng.integers(0, 376, N).astype(np.float32)
    ty = rng.integers(0, 376, M).astype(np.float32)
    ty[:n_true] = qy[:n_true] + rng.integers(-3, 4, n_true)  # true partners mostly inside the band, some just outside
    om, omap = O.match(q, t, mode=mode, qy=qy, ty=ty, band=band, num_threads=8)
    assert len(om) >= 3
    um, _ = O.match(q, t, mode=mode, num_threads=8)
    if band <= 2.0 and mode != 2:
        assert len(um) != len(om) or (um["trainIdx"] != om["trainIdx"]).any(), "the mask must change the result"
    fe = S.Frontend(0, 2, 64, 64, 16)
    for alg in (S.MATCHER_TENSOR, S.MATCHER_EXACT_FP32):
        gm, gmap = fe.match(q, t, mode=mode, algorithm=alg, q_kpts=_kp(S, qy), t_kpts=_kp(S, ty), band=band)
        assert len(gm) == len(om), (alg, len(gm), len(om))
        assert (gm["queryIdx"] == om["queryIdx"]).all() and (gm["trainIdx"] == om["trainIdx"]).all(), alg
        assert (gm["distance"].view(np.uint32) == om["distance"].view(np.uint32)).all(), alg
        assert (gmap == omap).all(), alg
    fe.close()


def test_masked_match_nothing_allowed_and_everything_allowed(spvo, oracle):
    from conftest import unit_rows
    S, O = spvo, oracle
    q, t = unit_rows(200, 1), unit_rows(260, 2)
    fe = S.Frontend(0, 2, 64, 64, 16)
    for alg in (S.MATCHER_TENSOR, S.MATCHER_EXACT_FP32):
        # disjoint rows: nothing is allowed
        gm, gmap = fe.match(q, t, mode=1, algorithm=alg, q_kpts=_kp(S, np.zeros(200)), t_kpts=_kp(S, np.full(260, 100.0)),
                            band=2.0)
        assert len(gm) == 0 and (gmap == -1).all()
        # a huge band = the unmasked matcher
        gm, gmap = fe.match(q, t, mode=1, algorithm=alg, q_kpts=_kp(S, np.zeros(200)), t_kpts=_kp(S, np.full(260, 100.0)),
                            band=1e6)
        om, omap = O.match(q, t, mode=1)
        assert len(gm) == len(om) and (gm["trainIdx"] == om["trainIdx"]).all() and (gmap == omap).all()
    fe.close()


@pytest.mark.parametrize("mode", [1, 2])
def test_stereo_pipeline_row_band_flag(spvo, oracle, mode):
    """SPVO_MATCH_FLAG_ROW_BAND in spvo_stereo_batch: L<->R problems masked with band = stereo_threshold, temporal ones
    unmasked; everything downstream (keep flags, quadruples, carry) consumes the masked lists."""
    import spvo_b200.synth as synth
    S, O = spvo, oracle
    H, W, K, F, NB = 192, 640, 500, 3, 2
    semi, desc = synth.make_stream(F * NB, H, W, seed=13, device="cpu")
    semi, desc = semi.numpy(), desc.numpy()
    thr, mind = 2.0, 0.25
    ref, prev, prev_maps = [], None, None
    for f in range(F * NB):
        d = O.decode(semi[f], desc[f], max_keypoints=K)
        nl, nr = int(d["n"][0]), int(d["n"][1])
        ms, maps = O.match(d["desc"][0, :nl], d["desc"][1, :nr], mode=mode, qy=d["kpts"]["y"][0, :nl],
                           ty=d["kpts"]["y"][1, :nr], band=thr)
        keep = O.stereo_filter(d["kpts"][0], d["kpts"][1], ms, thr, mind)
        if prev is not None:
            mt, mapt = O.match(d["desc"][0, :nl], prev["desc"][0, : int(prev["n"][0])], mode=mode)
            quads = O.consistency(ms, mapt, keep, prev_maps)
        else:
            mt, mapt, quads = np.zeros(0, O.DMATCH_DTYPE), np.full(nl, -1, np.int32), np.zeros((0, 4), np.int32)
        ref.append(dict(dec=d, ms=ms, maps=maps, keep=keep, mt=mt, mapt=mapt, quads=quads))
        prev, prev_maps = d, maps
    from test_gpu_bench_shape import _check_frame
    for alg in (S.MATCHER_TENSOR, S.MATCHER_EXACT_FP32):
        fe = S.Frontend(0, 2 * F, H, W, K)
        for b in range(NB):
            out = {k: v.numpy() for k, v in fe.alloc_stereo_out(F, K, device="cpu").items()}
            fe.stereo_batch(semi[b * F:(b + 1) * F], desc[b * F:(b + 1) * F], F, H, W, out, max_keypoints=K, mode=mode,
                            algorithm=alg, stereo_threshold=thr, min_disparity=mind, row_band=True)
            for f in range(F):
                _check_frame(S, out, ref[b * F + f], f, F, K, f"alg {alg} frame {b * F + f}")
        fe.close()
    assert all(r["keep"].all() or (np.abs(r["dec"]["kpts"]["x"][0][r["ms"]["queryIdx"]] - r["dec"]["kpts"]["x"][1][r["ms"]["trainIdx"]]) < mind).any() for r in ref), \
        "inside the band only the min-disparity test can still reject a masked match"
