"""result formats, and the downstream pose computed from the GPU lists equals the pose
computed from the oracle lists (north_star: <= 1e-6 relative translation; lists are bit-identical, so the
cv2 triangulation + PnP that the reference runs next (BASE:209-281) sees identical inputs)."""
import os

import numpy as np
import pytest


def test_kitti_pose_and_latency_formats(tmp_path, spvo):
    from spvo_b200 import formats as F
    T = np.eye(4)
    T[:3, 3] = [1.5, -0.25, 3.0]
    line = F.kitti_pose_line(T)
    assert len(line.split()) == 12 and line.split()[3] == "1.5" and line.split()[11] == "3.0"
    p = tmp_path / "00_pred.txt"
    assert F.write_kitti_poses(str(p), [T, T[:3]]) == 2
    back = F.read_kitti_poses(str(p))
    assert back.shape == (2, 3, 4) and np.array_equal(back[0], T[:3])
    assert F.latency_csv_row(7.39, 6.03, 1.63).split(",")[:3] == ["7.39", "6.03", "1.63"]
    assert abs(float(F.latency_csv_row(7.39, 6.03, 1.63).split(",")[3]) - 15.05) < 1e-9


def _pose_from_lists(cv2, kl, kr, kpl, ms, mt_map, P_l, P_r):
    """Reference flow at refinement_degree 0 (BASE:127-281): intersect maps, stereo checks, triangulate the
    previous pair is replaced by the current pair for this self-contained check, PnP-RANSAC."""
    pts_l, pts_r, pts_prev = [], [], []
    for m in ms:
        iL, iR = int(m["queryIdx"]), int(m["trainIdx"])
        iP = int(mt_map[iL])
        if iP < 0:
            continue
        a, b = kl[iL], kr[iR]
        if abs(a["y"] - b["y"]) > 2.0 or abs(a["x"] - b["x"]) < 0.25:   # BASE:169-172
            continue
        pts_l.append((a["x"], a["y"]))
        pts_r.append((b["x"], b["y"]))
        pts_prev.append((kpl[iP]["x"], kpl[iP]["y"]))
    n_pts = len(pts_l)
    pts_l, pts_r, pts_prev = (np.array(v, np.float64).T for v in (pts_l, pts_r, pts_prev))
    X = cv2.triangulatePoints(P_l, P_r, pts_l, pts_r)
    X = (X[:3] / X[3]).T
    ok, rvec, tvec, inl = cv2.solvePnPRansac(X, pts_prev.T.copy(), P_l[:, :3].copy(), None, iterationsCount=500,
                                            reprojectionError=2.0, confidence=0.999, flags=cv2.USAC_ACCURATE)
    return ok, rvec, tvec, n_pts


@pytest.mark.gpu
def test_pose_from_gpu_lists_equals_pose_from_oracle_lists(spvo, oracle):
    cv2 = pytest.importorskip("cv2")
    import spvo_b200.synth as synth
    H, W, K = 192, 640, 500
    semi, desc = synth.make_stream(2, H, W, seed=12, device="cpu")
    semi, desc = semi.numpy(), desc.numpy()
    fx = 400.0
    P_l = np.array([[fx, 0, W / 2, 0], [0, fx, H / 2, 0], [0, 0, 1, 0]], np.float64)
    P_r = P_l.copy()
    P_r[0, 3] = -fx * 0.54
    fe = spvo.Frontend(0, 2, H, W, K)
    poses = {}
    for name in ("gpu", "oracle"):
        dec = []
        for f in range(2):
            d = fe.decode(semi[f], desc[f], max_keypoints=K) if name == "gpu" else oracle.decode(semi[f], desc[f], max_keypoints=K)
            dec.append(d)
        d1, d0 = dec[1], dec[0]
        nl, nr, npl = int(d1["n"][0]), int(d1["n"][1]), int(d0["n"][0])
        mfun = fe.match if name == "gpu" else oracle.match
        ms, _ = mfun(d1["desc"][0, :nl], d1["desc"][1, :nr], mode=1)
        _, mt_map = mfun(d1["desc"][0, :nl], d0["desc"][0, :npl], mode=1)
        cv2.setRNGSeed(1)
        poses[name] = _pose_from_lists(cv2, d1["kpts"][0], d1["kpts"][1], d0["kpts"][0], ms, mt_map, P_l, P_r)
    (ok_g, r_g, t_g, n_g), (ok_o, r_o, t_o, n_o) = poses["gpu"], poses["oracle"]
    assert ok_g == ok_o and n_g == n_o and n_g > 50
    if ok_g:
        assert np.linalg.norm(t_g - t_o) <= 1e-6 * max(1.0, np.linalg.norm(t_o))
        assert np.linalg.norm(r_g - r_o) <= 1e-6
    fe.close()
