"""The multi-GPU path on a real GPU: sequence.run_sharded with a closure over Frontend.stereo_batch_device.

Frames are independent in the front end except for the one-frame temporal dependency (feature_detection.hpp:87-90),
so a sequence is split into contiguous frame ranges, one handle (= one rank / GPU) each; every rank first processes
its predecessor frame as a halo after a reset (clearLagecyData, feature_detection_base.cpp:35-66).  World sizes 2 and
3 are emulated with separate handles on cuda:0 (the ranks share nothing, so this is exactly what N processes do); the
per-frame lists must equal the single-handle sequential run, which other tests pin to the oracle.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _make_processor(S, fe, semi, desc, H, W, K, mode, batch):
    """process_batch(first_frame, count, reset) -> per-frame results, for sequence.run_shard."""
    import torch

    def process(first, count, reset):
        if reset:
            fe.stereo_reset()
        out = fe.alloc_stereo_out(count, K, device=semi.device)
        fe.stereo_batch_device(semi[first:first + count], desc[first:first + count], count, H, W, out,
                               max_keypoints=K, mode=mode, stereo_threshold=2.0, min_disparity=0.25)
        torch.cuda.synchronize()
        o = {k: v.cpu().numpy() for k, v in out.items()}
        kp = o["kpts"].view(S.KEYPOINT_DTYPE).reshape(2 * count, K)
        mm = o["matches"].view(S.DMATCH_DTYPE).reshape(2 * count, K)
        res = []
        for f in range(count):
            nl, nr = int(o["n_kpts"][2 * f]), int(o["n_kpts"][2 * f + 1])
            ns, nt, nq = int(o["n_matches"][f]), int(o["n_matches"][count + f]), int(o["n_quads"][f])
            res.append(dict(kl=kp[2 * f, :nl].copy(), kr=kp[2 * f + 1, :nr].copy(), ms=mm[f, :ns].copy(),
                            mt=mm[count + f, :nt].copy(), q2t_s=o["q2t"][f, :nl].copy(),
                            q2t_t=o["q2t"][count + f, :nl].copy(), keep=o["stereo_keep"][f, :ns].copy(),
                            quads=o["quads"][f, :nq].copy()))
        return res

    return process


@pytest.mark.parametrize("mode", [1, 2])
def test_sharded_with_halo_equals_sequential_on_gpu(spvo, mode):
    import torch
    import spvo_b200.sequence as seq
    import spvo_b200.synth as synth
    S = spvo
    H, W, K, NF, batch = 192, 640, 300, 41, 8
    dev = torch.device("cuda", 0)
    semi, desc = synth.make_stream(NF, H, W, seed=11, device=dev)

    fe = S.Frontend(0, 2 * batch, H, W, K)
    fe.set_stream(torch.cuda.current_stream().cuda_stream)
    ref = seq.run_shard(_make_processor(S, fe, semi, desc, H, W, K, mode, batch), 0, NF, batch)
    fe.close()
    assert len(ref) == NF and sum(len(r["mt"]) for r in ref) > 20 * NF and sum(len(r["quads"]) for r in ref) > 5 * NF

    for world in (2, 3):
        handles = [S.Frontend(0, 2 * batch, H, W, K) for _ in range(world)]
        parts = []
        for rank, h in enumerate(handles):
            h.set_stream(torch.cuda.current_stream().cuda_stream)
            parts.append(seq.run_sharded(_make_processor(S, h, semi, desc, H, W, K, mode, batch), NF, batch, rank,
                                         world))
        got = seq.run_sharded(lambda *a: [], 0, batch, 0, 1, gather=lambda mine: parts)  # gather in rank order
        for h in handles:
            h.close()
        assert len(got) == NF
        for f, (a, b) in enumerate(zip(got, ref)):
            for key in a:
                x, y = a[key], b[key]
                assert x.shape == y.shape and x.tobytes() == y.tobytes(), (world, f, key)


def test_shard_without_halo_loses_only_the_first_temporal_match(spvo):
    """halo=False is what a rank would get WITHOUT the predecessor frame: only its first frame differs (no temporal
    matches, no quadruples) -- documents why the halo frame exists."""
    import torch
    import spvo_b200.sequence as seq
    import spvo_b200.synth as synth
    S = spvo
    H, W, K, NF, batch = 128, 320, 200, 12, 4
    dev = torch.device("cuda", 0)
    semi, desc = synth.make_stream(NF, H, W, seed=4, device=dev)
    fe = S.Frontend(0, 2 * batch, H, W, K)
    fe.set_stream(torch.cuda.current_stream().cuda_stream)
    proc = _make_processor(S, fe, semi, desc, H, W, K, 1, batch)
    ref = seq.run_shard(proc, 0, NF, batch)
    first, count = seq.plan_shards(NF, 2)[1]
    nohalo = seq.run_shard(proc, first, count, batch, halo=False)
    halo = seq.run_shard(proc, first, count, batch, halo=True)
    fe.close()
    assert len(nohalo[0]["mt"]) == 0 and len(nohalo[0]["quads"]) == 0 and len(ref[first]["mt"]) > 0
    for i in range(count):
        for key in halo[i]:
            assert halo[i][key].tobytes() == ref[first + i][key].tobytes(), (i, key)
        if i > 1:  # frame first+1's quads use frame first's L<->R map, which does not depend on the halo
            for key in nohalo[i]:
                assert nohalo[i][key].tobytes() == ref[first + i][key].tobytes(), (i, key)
