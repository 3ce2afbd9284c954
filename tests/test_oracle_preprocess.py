"""CPU tests pinning the preprocessing oracle (crop + cv::resize INTER_LINEAR 8UC1 + /255 + projection-matrix patch,
reference feature_detection_base.cpp:68-121 and feature_detection_neural_network.cpp:139-161) to the reference's
real resize, cv::resize: live against cv2 when importable, and against the committed cv2-generated fixtures."""
import glob
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")
K255 = np.float32(1.0) / np.float32(255.0)


def test_golden_cv2_fixtures(oracle):
    files = sorted(glob.glob(os.path.join(GOLD, "preprocess_cv2_*.npz")))
    assert len(files) == 4
    for f in files:
        g = np.load(f)
        img, H, W = g["img"], int(g["H"]), int(g["W"])
        assert oracle.crop_geometry(img.shape[0], img.shape[1], H, W) == tuple(int(v) for v in g["crop"])
        inp, rs, _ = oracle.preprocess(img, H, W)
        assert (rs == g["resized"]).all(), f"{f}: resized image differs from cv2.resize"
        assert (inp.view(np.uint32) == (g["resized"].astype(np.float32) * K255).view(np.uint32)).all()


def test_live_cv2_bit_exact(oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(7)
    cases = [((375, 1242), (376, 1240)),   # KITTI image -> config 1/2 network input: 6-column crop, mild upscale
             ((370, 1226), (120, 392)),    # reference default ctor size
             ((480, 640), (192, 640)),     # row crop
             ((376, 1240), (376, 1240)),   # identity
             ((256, 512), (128, 256)),     # exact 2x decimation: OpenCV switches to INTER_AREA
             ((100, 300), (96, 96)), ((300, 100), (96, 96)), ((31, 17), (200, 120))]
    for (rows, cols), (H, W) in cases:
        img = rng.integers(0, 256, (rows, cols), dtype=np.uint8)
        cr, cc, ro, co = oracle.crop_geometry(rows, cols, H, W)
        ref = cv2.resize(img[ro:ro + cr, co:co + cc], (W, H), interpolation=cv2.INTER_LINEAR)
        inp, rs, _ = oracle.preprocess(img, H, W)
        assert (rs == ref).all(), ((rows, cols), (H, W), int((rs != ref).sum()))
        assert (inp == ref.astype(np.float32) * K255).all()
    # smooth images (long runs of equal neighbours) exercise the rounding of the fixed-point blend differently
    yy, xx = np.mgrid[0:375, 0:1242]
    img = ((np.sin(xx / 37.0) + np.cos(yy / 23.0)) * 60 + 128).astype(np.uint8)
    ref = cv2.resize(img[:, 3:1239], (1240, 376), interpolation=cv2.INTER_LINEAR)
    assert (oracle.preprocess(img, 376, 1240)[1] == ref).all()


def test_crop_geometry_follows_the_reference_arithmetic(oracle):
    # BASE:71-113: float aspect ratios, int <- float truncation, centred offsets
    assert oracle.crop_geometry(375, 1242, 376, 1240) == (375, 1236, 0, 3)
    assert oracle.crop_geometry(480, 640, 192, 640) == (192, 640, 144, 0)
    assert oracle.crop_geometry(376, 1240, 376, 1240) == (376, 1240, 0, 0)
    for rows, cols, H, W in [(375, 1242, 376, 1240), (370, 1226, 120, 392), (1080, 1920, 192, 640), (97, 131, 64, 96)]:
        real, exp = np.float32(cols) / np.float32(rows), np.float32(W) / np.float32(H)
        cr, cc, ro, co = rows, cols, 0, 0
        if exp > real:
            cr = int(np.float32(cols) / exp)
            ro = (rows - cr) // 2
        elif exp < real:
            cc = int(np.float32(rows) * exp)
            co = (cols - cc) // 2
        assert oracle.crop_geometry(rows, cols, H, W) == (cr, cc, ro, co)


def test_projection_matrix_patch(oracle):
    # KITTI seq 00 P0 / P1 (3x4 row-major); BASE:93/109: principal point minus the crop offset, BASE:119-120:
    # rows 0 and 1 times (float)W / (float)cropped_cols
    P = np.array([[718.856, 0.0, 607.1928, -386.1448], [0.0, 718.856, 185.2157, 0.0], [0.0, 0.0, 1.0, 0.0]], np.float32)
    img = np.zeros((375, 1242), np.uint8)
    _, _, Pp = oracle.preprocess(img, 376, 1240, P)
    exp = P.copy()
    exp[0, 2] -= np.float32(3)
    exp[:2] *= np.float32(1240) / np.float32(1236)
    assert (Pp.view(np.uint32) == exp.view(np.uint32)).all()
    _, _, Pp = oracle.preprocess(np.zeros((480, 640), np.uint8), 192, 640, P)
    exp = P.copy()
    exp[1, 2] -= np.float32(144)
    exp[:2] *= np.float32(640) / np.float32(640)
    assert (Pp.view(np.uint32) == exp.view(np.uint32)).all()


def test_live_cv2_random_geometries(oracle):
    """80 random (source size, network input size) pairs, up- and down-scaling by up to ~6x in either direction."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(2024)
    for _ in range(80):
        rows, cols = int(rng.integers(9, 400)), int(rng.integers(9, 600))
        H, W = int(rng.integers(8, 300)), int(rng.integers(8, 400))
        try:
            cr, cc, ro, co = oracle.crop_geometry(rows, cols, H, W)
        except ValueError:
            continue  # empty crop (extreme aspect ratios)
        img = rng.integers(0, 256, (rows, cols), dtype=np.uint8)
        ref = cv2.resize(img[ro:ro + cr, co:co + cc], (W, H), interpolation=cv2.INTER_LINEAR)
        rs = oracle.preprocess(img, H, W)[1]
        assert (rs == ref).all(), ((rows, cols), (H, W), (cr, cc, ro, co), int((rs != ref).sum()))
