"""Measures BASELINE.json configs 3-5 (bench.py is configs[1]); writes one JSON document.

  config 3  K = 2048, 1240x376, kNN-2 + 0.8 ratio test, 74 pairs per call           -> pairs/s
  config 4  640x192, K = 500, batch of ONE stereo pair, latency per call, plain launches and CUDA graph
  preprocess  296 KITTI-sized 8-bit images -> network input (crop + resize + /255), device-resident
  config 5  matching only, N = M in {256 ... 8192}, cross-check and ratio modes, tensor path vs exact
            fp32 path (and cv2.BFMatcher on the host when importable, as the reference's matcher)

usage: python scripts/bench_configs.py [out.json]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import spvo_b200 as S
import spvo_b200.synth as synth

dev = torch.device("cuda", 0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)


def timed(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters  # ms


out = {}

# ---------------- config 3 ----------------
H, W, K, F = 376, 1240, 2048, 74
semi, desc = synth.make_stream(2 * F, H, W, seed=1, device=dev)
semi = semi.view(2, F, 2, 65, H // 8, W // 8)
desc = desc.view(2, F, 2, 256, H // 8, W // 8)
fe = S.Frontend(0, 2 * F, H, W, K)
fe.set_stream(stream.cuda_stream)
o = fe.alloc_stereo_out(F, K, device=dev)
i = [0]


def step3():
    fe.stereo_batch_device(semi[i[0] % 2], desc[i[0] % 2], F, H, W, o, max_keypoints=K, mode=S.MATCH_KNN_RATIO)
    i[0] += 1


ms = timed(step3, 20)
out["config3_K2048_ratio"] = {"ms_per_step": ms, "pairs_per_step": F, "pairs_per_s": F / ms * 1e3,
                              "mean_keypoints": o["n_kpts"].float().mean().item(),
                              "mean_matches": o["n_matches"].float().mean().item(),
                              "fallback_rows_total": int(fe.debug_counters()[1])}
fe.close()
del semi, desc, o

# ---------------- config 4 ----------------
H, W, K, F = 192, 640, 500, 1
semi, desc = synth.make_stream(64, H, W, seed=2, device=dev)
fe = S.Frontend(0, 2, H, W, K)
fe.set_stream(stream.cuda_stream)
o = fe.alloc_stereo_out(F, K, device=dev)
i = [0]


def step4():
    fe.stereo_batch_device(semi[i[0] % 64], desc[i[0] % 64], F, H, W, o, max_keypoints=K, mode=S.MATCH_NN_CROSSCHECK)
    i[0] += 1


ms = timed(step4, 200, warm=10)
res4 = {"us_per_pair_stream_launches": ms * 1e3, "mean_keypoints": o["n_kpts"].float().mean().item(),
        "mean_matches": o["n_matches"].float().mean().item()}
try:  # the same call captured once into a CUDA graph (fixed input buffers), replayed
    g = torch.cuda.CUDAGraph()
    s0, d0 = semi[0].clone(), desc[0].clone()
    fe.stereo_batch_device(s0, d0, F, H, W, o, max_keypoints=K, mode=S.MATCH_NN_CROSSCHECK)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=stream):
        fe.stereo_batch_device(s0, d0, F, H, W, o, max_keypoints=K, mode=S.MATCH_NN_CROSSCHECK)
    res4["us_per_pair_cuda_graph"] = timed(lambda: g.replay(), 200, warm=10) * 1e3
except Exception as e:  # noqa: BLE001
    res4["cuda_graph_error"] = repr(e)[:200]
out["config4_640x192_K500_batch1"] = res4
fe.close()

# ---------------- config 5 ----------------
try:
    import cv2
except Exception:  # noqa: BLE001
    cv2 = None
sweep = {}
fe = S.Frontend(0, 1, 64, 64, 16)
fe.set_stream(stream.cuda_stream)
for N in (256, 512, 1024, 2048, 4096, 8192):
    base = synth.random_descriptors(N, seed=N, device=dev)
    q = base
    t = base + 0.05 * torch.randn(base.shape, device=dev)
    t = (t / t.norm(dim=1, keepdim=True))[torch.randperm(N, device=dev)].contiguous()
    mo = torch.zeros(N, 4, dtype=torch.int32, device=dev)
    nm = torch.zeros(1, dtype=torch.int32, device=dev)
    row = {}
    for mode, mname in ((S.MATCH_NN_CROSSCHECK, "crosscheck"), (S.MATCH_KNN_RATIO, "ratio")):
        for alg, aname in ((S.MATCHER_TENSOR, "tensor"), (S.MATCHER_EXACT_FP32, "exact_fp32")):
            if aname == "exact_fp32" and N > 4096:
                continue
            ms = timed(lambda: fe.match_device(q, N, t, N, mo, nm, None, mode=mode, algorithm=alg), 20)
            row[f"{mname}_{aname}_ms"] = ms
            row[f"{mname}_{aname}_matches"] = int(nm.item())
            if aname == "tensor":
                row[f"{mname}_tensor_tflops_algorithmic"] = 2.0 * N * N * 256 / (ms * 1e-3) / 1e12
    if cv2 is not None and N <= 4096:
        qh, th = q.cpu().numpy(), t.cpu().numpy()
        t0 = time.perf_counter()
        cv2.BFMatcher_create(cv2.NORM_L2, True).match(qh, th)
        row["cv2_crosscheck_ms"] = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        cv2.BFMatcher_create(cv2.NORM_L2, False).knnMatch(qh, th, 2)
        row["cv2_knn2_ms"] = (time.perf_counter() - t0) * 1e3
    sweep[str(N)] = row
out["config5_matching_sweep"] = sweep
out["config5_note"] = ("single problem per call (latency-bound below N~2048: one problem fills few SMs); "
                       "bench.py's step runs 148 such problems per launch")
fe.close()

# ---------------- fp16 network outputs (bench.py's workload, inputs as binary16) ----------------
H, W, K, F = 376, 1240, 1000, 148
semi, desc = synth.make_stream(2 * F, H, W, seed=0, device=dev)
semi = semi.view(2, F, 2, 65, H // 8, W // 8)
desc = desc.view(2, F, 2, 256, H // 8, W // 8)
s16, d16 = semi.half(), desc.half()
fe = S.Frontend(0, 2 * F, H, W, K)
fe.set_stream(stream.cuda_stream)
o = fe.alloc_stereo_out(F, K, device=dev)
i = [0]


def step16():
    fe.stereo_batch_device(s16[i[0] % 2], d16[i[0] % 2], F, H, W, o, max_keypoints=K, f16=True)
    i[0] += 1


def step32():
    fe.stereo_batch_device(semi[i[0] % 2], desc[i[0] % 2], F, H, W, o, max_keypoints=K)
    i[0] += 1


ms32 = timed(step32, 12)
fe.profile_enable(True)
ms16 = timed(step16, 12)
prof = fe.profile_read()
fe.profile_enable(False)
hs, hd = s16[0].cpu().pin_memory(), d16[0].cpu().pin_memory()
ho = fe.alloc_stereo_out(F, K, device="cpu", pinned=True)
fe.set_stream(0)
for _ in range(2):
    fe.stereo_batch(hs, hd, F, H, W, ho, max_keypoints=K, f16=True)
t0 = time.perf_counter()
for _ in range(5):
    fe.stereo_batch(hs, hd, F, H, W, ho, max_keypoints=K, f16=True)
e2e16 = 5 * F / (time.perf_counter() - t0)
out["fp16_inputs_148_pairs"] = {"ms_per_step_f16": ms16, "pairs_per_s_f16": F / ms16 * 1e3, "ms_per_step_f32_same_run": ms32,
                                "e2e_pairs_per_s_f16_host_buffers": e2e16,
                                "kernels_ms_f16": {k: v[0] / max(v[1], 1) for k, v in prof.items() if v[1]},
                                "note": "spvo_stereo_batch[_device]_f16: same results as the widened fp32 tensors; "
                                        "decode reads half the bytes, the host form moves half the bytes over PCIe"}
fe.close()
del semi, desc, s16, d16, o
torch.cuda.synchronize()
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)

# ---------------- preprocess (the step before the network) ----------------
rows, cols, H, W, B = 375, 1242, 376, 1240, 296
imgs = torch.randint(0, 256, (3, B, rows, cols), dtype=torch.uint8, device=dev)  # ring of 3 batches (412 MB > L2)
inp = torch.empty(B, H, W, dtype=torch.float32, device=dev)
rsz = torch.empty(B, H, W, dtype=torch.uint8, device=dev)
fe = S.Frontend(0, 2, 64, 64, 16)
fe.set_stream(stream.cuda_stream)
i = [0]


def stepp():
    fe.preprocess_device(imgs[i[0] % 3], B, rows, cols, cols, H, W, inp, rsz, None)
    i[0] += 1


ms = timed(stepp, 30)
alg_bytes = B * (rows * 1236 + H * W * 5)  # cropped 8-bit source read once + fp32 input and 8-bit image written
out["preprocess_375x1242_to_376x1240"] = {"images_per_call": B, "ms_per_call": ms, "images_per_s": B / ms * 1e3,
                                          "algorithmic_GBps": alg_bytes / (ms * 1e-3) / 1e9,
                                          "note": "crop + cv::resize INTER_LINEAR (bit-exact) + /255; bytes = crop read + fp32 and u8 outputs"}
fe.close()

txt = json.dumps(out, indent=1)
print(txt)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(txt + "\n")
