"""Where does the tensor-core matcher overtake the exact fp32 matcher?  (spvo_match_cfg.algorithm = AUTO is tuned from
this table: api.cu pick_algorithm.)  Two regimes:
  single   one problem per call (spvo_match_device), N = M
  batched  148 stereo pairs per call through spvo_stereo_batch_device (148 + 148 problems), K keypoints per image
usage: python scripts/match_sweep.py [out.json]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import spvo_b200 as S
import spvo_b200.synth as synth

dev = torch.device("cuda", 0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)


def timed(fn, iters, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


out = {"single_us": {}, "batched_us": {}}
fe = S.Frontend(0, 1, 64, 64, 16)
fe.set_stream(stream.cuda_stream)
for N in (32, 64, 96, 128, 192, 256, 320, 384, 448, 512, 640, 768, 1024, 1536, 2048):
    base = synth.random_descriptors(N, seed=N, device=dev)
    t = base + 0.05 * torch.randn(base.shape, device=dev)
    t = (t / t.norm(dim=1, keepdim=True))[torch.randperm(N, device=dev)].contiguous()
    mo = torch.zeros(N, 4, dtype=torch.int32, device=dev)
    nm = torch.zeros(1, dtype=torch.int32, device=dev)
    row = {}
    for mode, mname in ((S.MATCH_NN, "nn"), (S.MATCH_NN_CROSSCHECK, "crosscheck"), (S.MATCH_KNN_RATIO, "ratio")):
        for alg, aname in ((S.MATCHER_TENSOR, "tensor"), (S.MATCHER_EXACT_FP32, "exact")):
            row[f"{mname}_{aname}"] = round(timed(lambda: fe.match_device(base, N, t, N, mo, nm, None, mode=mode, algorithm=alg), 50), 2)
    out["single_us"][str(N)] = row
fe.close()

H, W, F = 376, 1240, 148
semi, desc = synth.make_stream(F, H, W, seed=0, device=dev)
for K in (64, 128, 192, 256, 384, 512, 768, 1000):
    fe = S.Frontend(0, 2 * F, H, W, K)
    fe.set_stream(stream.cuda_stream)
    o = fe.alloc_stereo_out(F, K, device=dev)
    row = {}
    for mode, mname in ((S.MATCH_NN_CROSSCHECK, "crosscheck"), (S.MATCH_KNN_RATIO, "ratio")):
        for alg, aname in ((S.MATCHER_TENSOR, "tensor"), (S.MATCHER_EXACT_FP32, "exact")):
            row[f"{mname}_{aname}"] = round(timed(lambda: fe.stereo_batch_device(semi, desc, F, H, W, o, max_keypoints=K, mode=mode, algorithm=alg), 10, warm=3), 1)
    out["batched_us"][str(K)] = row
    fe.close()
out["note"] = ("batched: whole stereo step (decode included, identical for both algorithms), so the difference of a "
               "row is the matcher's")
js = json.dumps(out, indent=1)
print(js)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(js + "\n")
