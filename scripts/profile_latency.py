"""One stereo pair per call at the latency configuration (640x192, K=500), for ncu captures:
   ncu ... python scripts/profile_latency.py [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import spvo_b200 as S
import spvo_b200.synth as synth

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
H, W, K = 192, 640, 500
semi, desc = synth.make_stream(reps, H, W, seed=0, device="cuda")
fe = S.Frontend(0, 2, H, W, K)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
fe.set_stream(st.cuda_stream)
out = fe.alloc_stereo_out(1, K, device="cuda")
for i in range(reps):
    fe.stereo_batch_device(semi[i:i + 1], desc[i:i + 1], 1, H, W, out, max_keypoints=K, mode=S.MATCH_NN_CROSSCHECK)
torch.cuda.synchronize()
print("kpts", out["n_kpts"].float().mean().item(), "matches", out["n_matches"].float().mean().item())
