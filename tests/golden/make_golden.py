"""Generates tests/golden/*.npz.  Run in the build container (needs cv2 for the matching fixtures):
    python tests/golden/make_golden.py
* match_cv2_*.npz   : inputs + outputs of the reference's real matcher, cv2.BFMatcher(NORM_L2)
                      (.match with / without crossCheck, .knnMatch(k=2) + the 0.8 ratio test of
                      feature_detection_base.cpp:466-472).  These PIN the matching oracle.
* preprocess_cv2_*.npz : small seeded 8-bit images + the output of cv2.resize(INTER_LINEAR) on the crop the
                      reference takes (feature_detection_base.cpp:68-121).  These PIN the preprocessing oracle.
* decode_oracle_*.npz : seeded inputs (regenerated from the seed) + SHA-256 of the oracle's decode
                      outputs.  The reference's decode cannot be run here (Eigen/OpenCV C++/ROS/TensorRT
                      absent), so these pin the oracle against regressions, not against the reference.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from conftest import make_inputs, unit_rows  # noqa: E402


def match_case(seed, N, M, noise):
    rng = np.random.default_rng(seed)
    base = unit_rows(max(N, M), seed)
    q = base[:N].copy()
    t = base + noise * rng.standard_normal(base.shape).astype(np.float32)
    t = (t / np.linalg.norm(t, axis=1, keepdims=True)).astype(np.float32)[rng.permutation(len(base))][:M]
    if M > 8 and N > 8:
        t[3] = t[1]          # duplicate train rows: lowest index must win
        q[5] = q[2]          # duplicate query rows: cross-check keeps the lower query
    return q, np.ascontiguousarray(t)


def cv2_outputs(q, t):
    import cv2
    out = {}
    for name, cc in (("nn", False), ("cc", True)):
        ms = cv2.BFMatcher_create(cv2.NORM_L2, crossCheck=cc).match(q, t)
        out[name + "_q"] = np.array([m.queryIdx for m in ms], np.int32)
        out[name + "_t"] = np.array([m.trainIdx for m in ms], np.int32)
        out[name + "_d"] = np.array([m.distance for m in ms], np.float32)
    knn = cv2.BFMatcher_create(cv2.NORM_L2, crossCheck=False).knnMatch(q, t, 2)
    keep = [m[0] for m in knn if len(m) == 2 and np.float32(m[0].distance) < np.float32(0.8) * np.float32(m[1].distance)]
    out["knn_q"] = np.array([m.queryIdx for m in keep], np.int32)
    out["knn_t"] = np.array([m.trainIdx for m in keep], np.int32)
    out["knn_d"] = np.array([m.distance for m in keep], np.float32)
    return out


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


DECODE_CASES = [  # (name, H, W, B, seed, sigma, K, conf, dist, border)
    ("small", 64, 96, 2, 1, 1.0, 200, 0.015, 4, 4),
    ("lowres", 192, 640, 1, 2, 1.0, 500, 0.015, 4, 4),
    ("ties", 128, 256, 1, 3, 0.1, 700, 0.015, 4, 4),
    ("wide_nms", 128, 256, 1, 4, 3.0, 300, 0.05, 8, 12),
]


def main():
    import cv2
    for i, (N, M, noise) in enumerate([(64, 80, 0.05), (200, 150, 0.2), (300, 300, 1.0)]):
        q, t = match_case(100 + i, N, M, noise)
        np.savez_compressed(os.path.join(HERE, f"match_cv2_{i}.npz"), q=q, t=t, cv2_version=cv2.__version__,
                            **cv2_outputs(q, t))
    from oracle import oracle as O
    # (rows, cols) -> (H, W): column crop + mild upscale (the KITTI case in small), row crop + downscale,
    # exact 2x decimation (OpenCV's INTER_AREA shortcut), upscale 2.1x
    for i, ((rows, cols), (H, W)) in enumerate([((47, 155), (48, 152)), ((120, 90), (24, 80)), ((64, 96), (32, 48)),
                                                ((30, 41), (64, 88))]):
        img = np.random.default_rng(500 + i).integers(0, 256, (rows, cols), dtype=np.uint8)
        cr, cc, ro, co = O.crop_geometry(rows, cols, H, W)
        ref = cv2.resize(img[ro:ro + cr, co:co + cc], (W, H), interpolation=cv2.INTER_LINEAR)
        np.savez_compressed(os.path.join(HERE, f"preprocess_cv2_{i}.npz"), img=img, H=H, W=W, crop=(cr, cc, ro, co),
                            resized=ref, cv2_version=cv2.__version__)
    for name, H, W, B, seed, sigma, K, conf, dist, border in DECODE_CASES:
        semi, desc = make_inputs(B, H, W, seed=seed, sigma=sigma)
        r = O.decode(semi, desc, conf_thresh=conf, dist_thresh=dist, border_remove=border, max_keypoints=K)
        np.savez_compressed(os.path.join(HERE, f"decode_oracle_{name}.npz"), n=r["n"],
                            kpts_xy=np.stack([r["kpts"]["x"], r["kpts"]["y"]], -1).astype(np.int16),
                            sha_kpts=sha(r["kpts"]), sha_desc=sha(r["desc"]), sha_scores=sha(r["scores"]),
                            input_sha=sha(semi, desc))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
