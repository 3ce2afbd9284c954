"""Per-phase clock stamps of k_detect from a DIAGNOSTIC build (-DSPVO_PHASE_TIMING, scripts/_diag/libspvo_timing.so):
   python scripts/detect_phases.py [H W K F]
Build: nvcc ... -DSPVO_PHASE_TIMING for every csrc/*.cu, link to scripts/_diag/libspvo_timing.so (see profiles/README.md)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import spvo_b200._lib as L

diag = os.path.join(ROOT, "scripts", "_diag", "libspvo_timing.so")
L.LIB_PATH = diag
import spvo_b200 as S
import spvo_b200.synth as synth

H, W, K, F = (int(a) for a in sys.argv[1:5]) if len(sys.argv) > 4 else (192, 640, 500, 1)
real = len(sys.argv) > 5
if real:
    g = np.load(os.path.join(ROOT, "tests", "golden", "realistic_kitti_1240x376.npz"))
    semi = torch.from_numpy(g["semi"].astype(np.float32)).cuda()[: 2 * F].reshape(F, 2, 65, H // 8, W // 8).contiguous()
    g = {"desc": np.repeat(g["desc"], 2, axis=0)[: 2 * F]} if g["desc"].shape[0] < 2 * F else g
    desc = torch.from_numpy(g["desc"].astype(np.float32)).cuda()[: 2 * F].reshape(F, 2, 256, H // 8, W // 8).contiguous()
else:
    semi, desc = synth.make_stream(F, H, W, seed=0, device="cuda")
fe = S.Frontend(0, 2 * F, H, W, K)
out = fe.alloc_stereo_out(F, K, device="cuda")
for _ in range(4):
    fe.stereo_batch_device(semi, desc, F, H, W, out, max_keypoints=K, mode=S.MATCH_NN_CROSSCHECK)
torch.cuda.synchronize()
lib = ctypes.CDLL(diag)
buf = (ctypes.c_longlong * (64 * 32))()
assert lib.spvo_debug_phase_clocks(buf) == 0
a = np.array(buf[:], dtype=np.int64).reshape(64, 32)
names = ["between", "G1 cell-max histogram", "G2 gather", "chunk select+sort", "A hash", "B NMS rounds", "C emit", "D bitmap",
         "tail", "outputs"]
for b in range(min(2 * F, 4)):
    tot = int(a[b, :12].sum() + a[b, 14] + a[b, 15])
    print(f"image {b}: total {tot} clk, {int(a[b, 12])} chunks, {int(a[b, 13])} NMS rounds:  " +
          "  ".join(f"{n}: {int(x)}" for n, x in zip(names, a[b, :10])) + f"  [NMS first rounds: {int(a[b, 10])}, later rounds: {int(a[b, 11])}; gather: stored cells {int(a[b, 14])}, recomputed cells {int(a[b, 15])}]  chunk sizes {[int(x) for x in a[b, 16:23] if x]} emitted after each {[int(x) for x in a[b, 24:31] if x]}")
