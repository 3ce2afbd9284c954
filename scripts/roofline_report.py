"""Per-kernel roofline table from a bench.py JSON line (live CUDA-event times) and the ncu traffic capture.

usage: python scripts/roofline_report.py [profiles/bench_r01.json] [profiles/ncu_traffic.json] > profiles/roofline_r01.txt

Columns: ms per step (events), share of the summed kernel time, algorithmic work per launch (bytes for the HBM-bound
kernels, flops for the GEMM), achieved rate, fraction of the measured peak
(MEASURED_PEAKS.json, else the profiling guide's fallback), the DRAM traffic ncu saw for one launch and that
traffic over the live launch time (above the algorithmic rate = intermediates / re-reads; below = L2 hits).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
bench = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "bench_r01.json")
traffic = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "ncu_traffic.json")
d = json.loads([l for l in open(bench).read().splitlines() if l.startswith("{")][-1])
tr = json.load(open(traffic)) if os.path.exists(traffic) else {}
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peaks = json.load(open(pk)) if os.path.exists(pk) else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
hbm, tf = peaks["hbm_gbs"], peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))

cfg = d["config"]
F, H, W, K = cfg["pairs_per_step"], cfg["H"], cfg["W"], cfg["keypoints"]
B, cells = 2 * F, (H // 8) * (W // 8)
n_kp = cfg.get("mean_keypoints", K)
alg = {  # (bound, algorithmic work per launch)
    "k_softmax_heat": ("hbm", B * 4 * 65 * cells),
    "k_detect": ("hbm", B * K * 28),
    "k_desc_planes": ("hbm", B * 4 * 256 * cells),
    "k_desc_normalize": ("hbm", B * K * 1024),
    "k_tc_gemm": ("tensor", 2 * F * 2.0 * n_kp * n_kp * 256),
    "k_tc_fill_dist": ("hbm", 2 * F * cfg.get("mean_matches", K) * 2048),
}
ks = d["kernels"]
tot = sum(v["ms_per_launch"] * v["launches"] for v in ks.values())
print(f"# {os.path.relpath(bench, ROOT)}: {d['value']:.0f} {d['unit']}, {d['ms_per_step']:.4f} ms per {F}-pair step, "
      f"{d['n_gpus']} GPU(s); peaks: HBM {hbm} GB/s, bf16 {tf} TFLOP/s (sustained)")
print(f"{'kernel':20s} {'ms/step':>8s} {'share':>6s} {'bound':>7s} {'algorithmic':>13s} {'achieved':>14s} {'of peak':>8s} {'ncu DRAM/launch':>16s} {'DRAM rate':>10s}")
for k, v in sorted(ks.items(), key=lambda kv: -kv[1]["ms_per_launch"] * kv[1]["launches"]):
    ms = v["ms_per_launch"] * v["launches"] / d["steps"]
    share = v["ms_per_launch"] * v["launches"] / tot
    bound, work = alg.get(k, ("-", 0))
    per_launch_s = v["ms_per_launch"] * 1e-3
    if bound == "hbm":
        ach, frac, w = f"{work / per_launch_s / 1e9:8.0f} GB/s", work / per_launch_s / 1e9 / hbm, f"{work / 1e6:9.1f} MB"
    elif bound == "tensor":
        ach, frac, w = f"{work / per_launch_s / 1e12:6.0f} TFLOP/s", work / per_launch_s / 1e12 / tf, f"{work / 1e9:7.1f} GFLOP"
    else:
        ach, frac, w = "-", None, "-"
    t = tr.get(f"{k}@F{F}")
    print(f"{k:20s} {ms:8.4f} {share:6.3f} {bound:>7s} {w:>13s} {ach:>14s} {('%.2f' % frac) if frac is not None else '-':>8s} "
          f"{('%.1f MB' % (t / 1e6)) if t else '-':>16s} {('%.0f GB/s' % (t / per_launch_s / 1e9)) if t else '-':>10s}")
if "decode_roofline" in d:
    r = d["decode_roofline"]
    print(f"all decode kernels: {r['achieved']:.0f} GB/s algorithmic = {r['frac']:.2f} of the HBM peak")
print(f"end to end (host buffers): {d['e2e']['value']:.0f} {d['e2e']['unit']}; H2D {d['e2e']['h2d_bytes_per_step'] / 1e9:.2f} GB per step")
