"""One decode of the realistic fixtures replicated to a 296-image batch, for ncu captures of k_detect:
   ncu ... python scripts/profile_realistic.py [K]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import spvo_b200 as S

K = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
H, W, B = 376, 1240, 296
gold = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                            "realistic_kitti_1240x376.npz"))
dev = torch.device("cuda", 0)
semi = torch.from_numpy(gold["semi"].astype(np.float32)).to(dev).repeat(B // 4, 1, 1, 1).contiguous()
fe = S.Frontend(0, B, H, W, K)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
fe.set_stream(st.cuda_stream)
kp = torch.zeros(B, K, 7, device=dev)
n = torch.zeros(B, dtype=torch.int32, device=dev)
for _ in range(2):
    fe.decode_device(semi, None, B, H, W, kp, None, n, max_keypoints=K)
torch.cuda.synchronize()
print("mean keypoints", n.float().mean().item(), "slow images", fe.debug_counters()[0])
