"""CPU tests pinning the matching oracle to the reference's real matcher, cv::BFMatcher (OpenCV):
live against cv2 when it is importable, and against the committed cv2-generated fixtures."""
import glob
import os

import numpy as np
import pytest

from conftest import unit_rows

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _same(m, q, t, d):
    assert len(m) == len(q)
    assert (m["queryIdx"] == q).all() and (m["trainIdx"] == t).all()
    assert (m["distance"].view(np.uint32) == d.view(np.uint32)).all(), "distance bits differ from cv2"


def test_golden_cv2_fixtures(oracle):
    files = sorted(glob.glob(os.path.join(GOLD, "match_cv2_*.npz")))
    assert len(files) == 3
    for f in files:
        g = np.load(f)
        q, t = g["q"], g["t"]
        for mode, key in ((oracle.MODE_NN, "nn"), (oracle.MODE_NN_CROSSCHECK, "cc"), (oracle.MODE_KNN_RATIO, "knn")):
            m, q2t = oracle.match(q, t, mode=mode, ratio=0.8)
            _same(m, g[key + "_q"], g[key + "_t"], g[key + "_d"])
            exp = np.full(len(q), -1, np.int32)
            exp[g[key + "_q"]] = g[key + "_t"]
            assert (q2t == exp).all()
            assert (np.diff(m["queryIdx"]) > 0).all()  # ascending queryIdx, as solveStereoOdometry walks it


def test_live_cv2_bit_exact(oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for N, M in ((150, 170), (64, 1), (1, 64), (257, 255)):
        q, t = unit_rows(N, N), unit_rows(M, M + 1)
        if M > 10:
            t[5] = t[9]
            t[7] = q[min(3, N - 1)]
        for cc in (False, True):
            ms = cv2.BFMatcher_create(cv2.NORM_L2, crossCheck=cc).match(q, t)
            m, _ = oracle.match(q, t, mode=1 if cc else 0)
            _same(m, np.array([x.queryIdx for x in ms], np.int32), np.array([x.trainIdx for x in ms], np.int32),
                  np.array([x.distance for x in ms], np.float32))
        if M >= 2:
            knn = cv2.BFMatcher_create(cv2.NORM_L2, False).knnMatch(q, t, 2)
            keep = [x[0] for x in knn if np.float32(x[0].distance) < np.float32(0.8) * np.float32(x[1].distance)]
            m, _ = oracle.match(q, t, mode=2, ratio=0.8)
            assert [x.queryIdx for x in keep] == list(m["queryIdx"]) and [x.trainIdx for x in keep] == list(m["trainIdx"])
    # non-unit-norm data: the lane order of hal::normL2Sqr_ matters more
    q = (rng.standard_normal((40, 256)) * 7).astype(np.float32)
    t = (rng.standard_normal((50, 256)) * 7).astype(np.float32)
    ms = cv2.BFMatcher_create(cv2.NORM_L2, False).match(q, t)
    m, _ = oracle.match(q, t, mode=0)
    _same(m, np.array([x.queryIdx for x in ms], np.int32), np.array([x.trainIdx for x in ms], np.int32),
          np.array([x.distance for x in ms], np.float32))


def test_distance_is_bitwise_symmetric(oracle):
    a, b = unit_rows(20, 1), unit_rows(20, 2)
    for i in range(20):
        assert np.float32(oracle.l2dist(a[i], b[i])).view(np.uint32) == np.float32(oracle.l2dist(b[i], a[i])).view(np.uint32)


def test_defined_edge_cases(oracle):
    q, t = unit_rows(5), unit_rows(6, 1)
    for mode in (0, 1, 2):
        m, q2t = oracle.match(np.zeros((0, 256), np.float32), t, mode=mode)
        assert len(m) == 0 and len(q2t) == 0
        m, q2t = oracle.match(q, np.zeros((0, 256), np.float32), mode=mode)
        assert len(m) == 0 and (q2t == -1).all()
    m, _ = oracle.match(q, t[:1], mode=2)     # reference: UB at BASE:469; defined here as "no match"
    assert len(m) == 0
    m, _ = oracle.match(q, t[:1], mode=0)
    assert len(m) == 5 and (m["trainIdx"] == 0).all()


def test_ratio_boundary_is_strict(oracle):
    """keep iff d0 < 0.8f * d1 in fp32 (BASE:469): construct d0 == 0.8f*d1 exactly -> rejected."""
    q = np.zeros((1, 256), np.float32)
    t = np.zeros((2, 256), np.float32)
    t[0, 0] = 4.0   # d0 = 4
    t[1, 0] = 5.0   # d1 = 5: 0.8f * 5 rounds to exactly 4.0f in fp32 -> 4 < 4 is false: rejected
    m, _ = oracle.match(q, t, mode=2, ratio=0.8)
    assert len(m) == 0
    t[1, 0] = np.nextafter(np.float32(5.0), np.float32(6.0))   # one ulp above the boundary
    t[1, 0] = 5.000002
    m, _ = oracle.match(q, t, mode=2, ratio=0.8)
    assert len(m) == 1 and m[0]["trainIdx"] == 0 and m[0]["distance"] == 4.0
    m, _ = oracle.match(q, t[[0, 0]], mode=2, ratio=1.0)   # d0 == d1 with ratio 1: strict '<' -> rejected
    assert len(m) == 0


def test_stereo_filter(oracle):
    kl = np.zeros(3, oracle.KEYPOINT_DTYPE)
    kr = np.zeros(3, oracle.KEYPOINT_DTYPE)
    kl["x"], kl["y"] = [100, 100, 100], [50, 50, 50]
    kr["x"], kr["y"] = [90, 99.9, 90], [51, 50, 52.5]
    m = np.zeros(3, oracle.DMATCH_DTYPE)
    m["queryIdx"] = m["trainIdx"] = [0, 1, 2]
    keep = oracle.stereo_filter(kl, kr, m, 2.0, 0.25)   # BASE:169-172
    assert list(keep) == [True, False, False]


def _cv2_masked(q, t, qy, ty, band, mode, ratio=0.8):
    """The masked matching AS cv2 does it: two masked BFMatcher(NORM_L2, crossCheck=False) calls + a manual mutual test
    (cv::BFMatcher rejects crossCheck with a mask), or knnMatch(k=2, mask) + the reference's ratio test."""
    import cv2
    mask = (np.abs(qy[:, None] - ty[None, :]) <= band).astype(np.uint8)
    bf = cv2.BFMatcher(cv2.NORM_L2, False)
    if mode == 2:
        out = []
        for m in bf.knnMatch(q, t, 2, mask=mask):
            if len(m) == 2 and m[0].distance < np.float32(ratio) * np.float32(m[1].distance):
                out.append((m[0].queryIdx, m[0].trainIdx, m[0].distance))
        return out
    fwd = bf.match(q, t, mask=mask)
    if mode == 0:
        return [(m.queryIdx, m.trainIdx, m.distance) for m in fwd]
    rev = {m.queryIdx: m.trainIdx for m in bf.match(t, q, mask=np.ascontiguousarray(mask.T))}
    return [(m.queryIdx, m.trainIdx, m.distance) for m in fwd if rev.get(m.trainIdx, -1) == m.queryIdx]


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_masked_matching_equals_masked_cv2(oracle, mode):
    """Row-band mask (north star): the oracle's masked matching is what masked cv2 calls give."""
    cv2 = pytest.importorskip("cv2")
    from conftest import unit_rows
    rng = np.random.default_rng(5)
    for N, M, band in ((300, 280, 2.0), (64, 200, 0.5), (500, 500, 8.0), (40, 3, 1.0)):
        q, t = unit_rows(N, seed=N), unit_rows(M, seed=M + 1)
        t[: min(N, M) // 2] = q[: min(N, M) // 2] + 0.05 * rng.standard_normal((min(N, M) // 2, 256)).astype(np.float32)
        qy = rng.integers(0, 60, N).astype(np.float32)
        ty = (qy[rng.integers(0, N, M)] + rng.integers(-3, 4, M)).astype(np.float32)
        want = _cv2_masked(q, t, qy, ty, band, mode)
        got, q2t = oracle.match(q, t, mode=mode, qy=qy, ty=ty, band=band)
        assert [(int(m["queryIdx"]), int(m["trainIdx"])) for m in got] == [(a, b) for a, b, _ in want]
        assert [np.float32(m["distance"]).view(np.uint32) for m in got] == [np.float32(d).view(np.uint32) for _, _, d in want]
        assert all(q2t[a] == b for a, b, _ in want) and (q2t >= 0).sum() == len(want)
    # no mask arguments = the unmasked matcher
    a, _ = oracle.match(q, t, mode=mode)
    b, _ = oracle.match(q, t, mode=mode, qy=qy, ty=ty, band=-1.0)
    assert a.tobytes() == b.tobytes()
