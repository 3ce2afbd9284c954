"""CPU tests of the decode oracle: hand-computable known answers, an independent
numpy restatement of the softmax / heatmap / greedy NMS, and the committed golden hashes."""
import glob
import hashlib
import os

import numpy as np
import pytest

from conftest import make_inputs

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def one_hot(H, W, points, amp=12.0):
    semi = np.zeros((1, 65, H // 8, W // 8), np.float32)
    semi[:, 64] = 8.0
    for k, (x, y) in enumerate(points):
        semi[0, (y % 8) * 8 + (x % 8), y // 8, x // 8] = amp - 0.5 * k
    return semi


def test_exp_matches_libm_within_2ulp(oracle):
    x = np.linspace(-20, 20, 4001).astype(np.float32)
    e = oracle.exp(x)
    ref = np.exp(x.astype(np.float64))
    assert np.max(np.abs(e - ref) / ref) < 2.5e-7
    assert oracle.exp(np.float32(0.0))[0] == 1.0


def test_heatmap_is_softmax_depth_to_space(oracle):
    semi, _ = make_inputs(2, 64, 96, seed=7)
    heat = oracle.heatmap(semi)
    e = np.exp(semi.astype(np.float64))
    p = e / (e.sum(1, keepdims=True) + 1e-5)
    nodust = p[:, :64].transpose(0, 2, 3, 1).reshape(2, 8, 12, 8, 8).transpose(0, 1, 3, 2, 4).reshape(2, 64, 96)
    assert np.abs(heat - nodust).max() < 1e-6
    # heat[8hc+i, 8wc+j] = p[8i+j, hc, wc]  (NN:289-326)
    assert abs(heat[1, 8 * 3 + 5, 8 * 7 + 2] - p[1, 8 * 5 + 2, 3, 7]) < 1e-6


def test_known_answers(oracle):
    H, W = 64, 96
    _, desc = make_inputs(1, H, W)
    r = oracle.decode(one_hot(H, W, [(41, 27)]), desc, max_keypoints=50)
    assert r["n"][0] == 1
    kp = r["kpts"][0, 0]
    assert (kp["x"], kp["y"], kp["size"], kp["angle"], kp["response"], kp["octave"], kp["class_id"]) == (41, 27, 1, -1, 0, 0, -1)
    assert oracle.decode(one_hot(H, W, [(40, 30), (44, 30)]), desc, max_keypoints=50)["n"][0] == 1   # Chebyshev 4: suppressed
    r = oracle.decode(one_hot(H, W, [(40, 30), (45, 30)]), desc, max_keypoints=50)                    # Chebyshev 5: both
    assert r["n"][0] == 2 and list(r["kpts"][0, :2]["x"]) == [40, 45]
    assert oracle.decode(one_hot(H, W, [(2, 30), (5, 30)]), desc, max_keypoints=50)["n"][0] == 0      # border suppresses, not emitted
    r = oracle.decode(one_hot(H, W, [(10 + 8 * i, 20) for i in range(8)]), desc, max_keypoints=3)     # K cut
    assert r["n"][0] == 3 and list(r["kpts"][0]["x"]) == [10, 18, 26]
    # strict '>' threshold: constant logits give 1/65 = 0.01538 everywhere
    z = np.zeros((1, 65, 8, 12), np.float32)
    assert oracle.decode(z, desc, conf_thresh=0.0153, max_keypoints=5)["n"][0] == 5
    assert oracle.decode(z, desc, conf_thresh=float(np.float32(1.0) / np.float32(65.0)) + 1e-4, max_keypoints=5)["n"][0] == 0


def test_tie_order_is_column_major(oracle):
    """Equal scores: canonical order = x ascending, then y ascending (stable sort of NN:205-217's list)."""
    H, W = 64, 96
    _, desc = make_inputs(1, H, W)
    z = np.zeros((1, 65, 8, 12), np.float32)
    r = oracle.decode(z, desc, conf_thresh=0.015, dist_thresh=4, border_remove=4, max_keypoints=30)
    xy = [(int(k["x"]), int(k["y"])) for k in r["kpts"][0, : r["n"][0]]]
    # first kept point is (0,0) [border, not emitted]; emitted points start at x=5 walking down the column
    assert xy[:3] == [(5, 5), (5, 10), (5, 15)]
    assert xy == sorted(xy)


def test_greedy_nms_against_python_restatement(oracle):
    H, W, K, d, b = 64, 96, 40, 4, 4
    semi, desc = make_inputs(1, H, W, seed=11)
    r = oracle.decode(semi, desc, max_keypoints=K, want_heat=True)
    heat = r["heat"][0]
    cands = [(heat[y, x], x, y) for x in range(W) for y in range(H) if heat[y, x] > np.float32(0.015)]
    cands.sort(key=lambda c: -c[0])  # python's sort is stable -> canonical order
    nms = np.zeros((H, W), bool)
    out = []
    for s, x, y in cands:
        if not nms[y, x]:
            if b <= y < H - b and b <= x < W - b:
                out.append((x, y))
            nms[max(0, y - d):y + d + 1, max(0, x - d):x + d + 1] = True
        if len(out) >= K:
            break
    got = [(int(k["x"]), int(k["y"])) for k in r["kpts"][0, : r["n"][0]]]
    assert got == out


def test_faithful_unstable_sort_differs_only_on_ties(oracle):
    semi, desc = make_inputs(1, 128, 256, seed=3, sigma=1.0)
    a = oracle.decode(semi, desc, max_keypoints=500)
    f = oracle.decode(semi, desc, max_keypoints=500, faithful_sort=True)
    n = int(a["n"][0])
    diff = np.nonzero(a["kpts"][0, :n] != f["kpts"][0, :n])[0]
    for i in diff:  # any disagreement must be between bit-equal scores
        assert a["scores"][0, i] == f["scores"][0, i]


def test_descriptor_sampling_properties(oracle):
    H, W = 64, 96
    semi, _ = make_inputs(1, H, W, seed=5)
    const = np.ones((1, 256, 8, 12), np.float32) * 0.3
    r = oracle.decode(semi, const, max_keypoints=50)
    n = int(r["n"][0])
    assert n > 10 and np.abs(r["desc"][0, :n] - 1.0 / 16.0).max() < 1e-6          # normalised constant = 1/sqrt(256)
    assert np.abs(np.linalg.norm(r["desc"][0, :n], axis=1) - 1).max() < 1e-6
    # a field linear in the cell index samples linearly in (x, y) under the align-corners map (NN:377-382)
    lin = np.zeros((1, 256, 8, 12), np.float32)
    rr, cc = np.meshgrid(np.arange(8), np.arange(12), indexing="ij")
    lin[0, 0], lin[0, 1], lin[0, 2] = rr, cc, 1.0
    o = oracle.decode(semi, lin, max_keypoints=50)
    for i in range(n):
        x, y = o["kpts"][0, i]["x"], o["kpts"][0, i]["y"]
        v = o["desc"][0, i]
        assert abs(v[0] / v[2] - y / (H - 1) * 7) < 1e-4 and abs(v[1] / v[2] - x / (W - 1) * 11) < 1e-4


def _sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def test_golden_decode_hashes(oracle):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    assert len(glob.glob(os.path.join(GOLD, "decode_oracle_*.npz"))) == len(mg.DECODE_CASES)
    for name, H, W, B, seed, sigma, K, conf, dist, border in mg.DECODE_CASES:
        g = np.load(os.path.join(GOLD, f"decode_oracle_{name}.npz"))
        semi, desc = make_inputs(B, H, W, seed=seed, sigma=sigma)
        assert _sha(semi, desc) == str(g["input_sha"]), "numpy RNG stream changed: regenerate the fixtures"
        r = oracle.decode(semi, desc, conf_thresh=conf, dist_thresh=dist, border_remove=border, max_keypoints=K)
        assert (r["n"] == g["n"]).all()
        assert _sha(r["kpts"]) == str(g["sha_kpts"]) and _sha(r["desc"]) == str(g["sha_desc"])
        assert _sha(r["scores"]) == str(g["sha_scores"])
