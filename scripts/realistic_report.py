"""Decode cost on REALISTIC network outputs vs the synthetic N(0,1) logits bench.py uses.

Replicates the committed fixtures (tests/golden/realistic_kitti_1240x376.npz: the reference's sp_mbv1 model on its
own KITTI sample images) to a 296-image batch and reports, per kernel, the live CUDA-event time of
spvo_decode_device, plus debug_counters[0] (images that needed more than the first candidate chunk).

usage: python scripts/realistic_report.py [out.json]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import spvo_b200 as S
import spvo_b200.synth as synth

dev = torch.device("cuda", 0)
H, W, B = 376, 1240, 296
gold = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                            "realistic_kitti_1240x376.npz"))
rs = torch.from_numpy(gold["semi"].astype(np.float32)).to(dev)   # [4,65,47,155]
rd = torch.from_numpy(gold["desc"].astype(np.float32)).to(dev)   # [2,256,47,155]
real_semi = rs.repeat(B // 4, 1, 1, 1).contiguous()
real_desc = rd.repeat(B // 2, 1, 1, 1).contiguous()
syn_semi, syn_desc = synth.make_stream(B // 2, H, W, seed=0, device=dev)
syn_semi, syn_desc = syn_semi.view(B, 65, H // 8, W // 8), syn_desc.view(B, 256, H // 8, W // 8)

stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
out = {}
for K in (1000, 2048):
    fe = S.Frontend(0, B, H, W, K)
    fe.set_stream(stream.cuda_stream)
    kp = torch.zeros(B, K, 7, device=dev)
    de = torch.zeros(B, K, 256, device=dev)
    n = torch.zeros(B, dtype=torch.int32, device=dev)
    for name, (se, ds) in (("synthetic", (syn_semi, syn_desc)), ("realistic", (real_semi, real_desc))):
        c0 = fe.debug_counters()[0]
        for _ in range(3):
            fe.decode_device(se, ds, B, H, W, kp, de, n, max_keypoints=K)
        fe.sync()
        fe.profile_enable(True)
        fe.profile_read()
        iters = 10
        for _ in range(iters):
            fe.decode_device(se, ds, B, H, W, kp, de, n, max_keypoints=K)
        prof = fe.profile_read()
        fe.profile_enable(False)
        c1 = fe.debug_counters()[0]
        out[f"K{K}_{name}"] = {
            "kernels_ms": {k: v[0] / iters for k, v in prof.items()},
            "decode_ms": sum(v[0] for v in prof.values()) / iters,
            "mean_keypoints": n.float().mean().item(),
            "slow_path_images_per_call": (c1 - c0) / (iters + 3),
        }
    fe.close()
print(json.dumps(out, indent=1))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
