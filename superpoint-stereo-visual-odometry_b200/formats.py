"""Result / wire formats of the reference's evaluation flow, so outputs can be
compared with the reference's tooling.  Host-side text formats only -- nothing here is on the hot path.

* KITTI odometry pose file (`kitti_results/<desc>/NN_pred.txt`): one line per frame, the row-major 3x4
  matrix [R|t] of cam0_start_T_cam0_curr, 12 numbers separated by single spaces
  (src/odml_data_processing/src/data_processing_node.cpp:175-187).
* latency CSV (`kitti_latency_csvs/<machine>/<cfg>_seq_<id>.csv`): one row per frame with the four
  millisecond columns feature detection, matching, solving, total
  (src/odml_visual_odometry/src/visual_odometry_node.cpp:246-258).
"""
from __future__ import annotations

from typing import Iterable, Sequence

import numpy as np


def kitti_pose_line(T: np.ndarray) -> str:
    """3x4 (or 4x4) pose matrix -> one KITTI odometry line (12 numbers, row-major)."""
    T = np.asarray(T, dtype=np.float64)
    if T.shape not in ((3, 4), (4, 4)):
        raise ValueError("pose must be 3x4 or 4x4")
    return " ".join(repr(float(v)) for v in T[:3, :4].reshape(-1))


def write_kitti_poses(path: str, poses: Iterable[np.ndarray]) -> int:
    n = 0
    with open(path, "w") as f:
        for T in poses:
            f.write(kitti_pose_line(T) + "\n")
            n += 1
    return n


def read_kitti_poses(path: str) -> np.ndarray:
    rows = [np.array(ln.split(), dtype=np.float64).reshape(3, 4) for ln in open(path) if ln.strip()]
    return np.stack(rows) if rows else np.zeros((0, 3, 4))


def latency_csv_row(t_detect_ms: float, t_match_ms: float, t_solve_ms: float) -> str:
    """One row of the reference's latency CSV: detection, matching, solving, total (ms)."""
    return f"{t_detect_ms},{t_match_ms},{t_solve_ms},{t_detect_ms + t_match_ms + t_solve_ms}"


def write_latency_csv(path: str, rows: Sequence[Sequence[float]]) -> None:
    with open(path, "w") as f:
        for r in rows:
            f.write(latency_csv_row(*r) + "\n")
