"""One stereo batch (the bench workload) for ncu captures:  ncu ... python scripts/profile_step.py [F] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import spvo_b200 as S
import spvo_b200.synth as synth

F = int(sys.argv[1]) if len(sys.argv) > 1 else 148
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
alg = int(sys.argv[3]) if len(sys.argv) > 3 else 0
H, W, K = 376, 1240, 1000
semi, desc = synth.make_stream(F, H, W, seed=0, device="cuda")
fe = S.Frontend(0, 2 * F, H, W, K)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
fe.set_stream(st.cuda_stream)
out = fe.alloc_stereo_out(F, K, device="cuda")
for _ in range(reps):
    fe.stereo_batch_device(semi, desc, F, H, W, out, max_keypoints=K, mode=S.MATCH_NN_CROSSCHECK, algorithm=alg)
torch.cuda.synchronize()
print("counters", fe.debug_counters()[:3], "kpts", out["n_kpts"].float().mean().item(), "matches",
      out["n_matches"].float().mean().item())
