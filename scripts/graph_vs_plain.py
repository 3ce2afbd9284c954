"""148 stereo pairs per call with plain launches vs the library-owned CUDA graph (spvo_set_graph_mode), alternating:
   python scripts/graph_vs_plain.py      (B200: 1.117 vs 1.090 ms per step)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import spvo_b200 as S
import spvo_b200.synth as synth
F, H, W, K, R = 148, 376, 1240, 1000, 3
semi, desc = synth.make_stream(R * F, H, W, seed=0, device="cuda")
semi = semi.view(R, F, 2, 65, H // 8, W // 8); desc = desc.view(R, F, 2, 256, H // 8, W // 8)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
for graph in (False, True, False, True):
    fe = S.Frontend(0, 2 * F, H, W, K); fe.set_stream(st.cuda_stream); fe.set_graph_mode(graph)
    out = fe.alloc_stereo_out(F, K, device="cuda")
    for i in range(14):
        fe.stereo_batch_device(semi[i % R], desc[i % R], F, H, W, out, max_keypoints=K, mode=S.MATCH_NN_CROSSCHECK)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for i in range(60):
        fe.stereo_batch_device(semi[i % R], desc[i % R], F, H, W, out, max_keypoints=K, mode=S.MATCH_NN_CROSSCHECK)
    e1.record(st); torch.cuda.synchronize()
    print("graph" if graph else "plain", e0.elapsed_time(e1) / 60, "ms/step", flush=True)
    fe.close()
