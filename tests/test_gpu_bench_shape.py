"""Oracle parity of the stereo pipeline AT THE SHAPES bench.py MEASURES (BASELINE.json configs 2 and 3).

The pipeline glue -- operand slots written by k_desc_normalize, the carry between batches, padded slot
capacities, the proved fp16 bound of the tensor matcher, the 4-chunk host path -- is exercised here at
1240x376 with K = 1000 / F = 148 (NN + cross-check) and K = 2048 (kNN ratio 0.8), and sampled frames are
compared bit for bit with the CPU oracle run the way the reference's stereoCallback runs
(visual_odometry_node.cpp:175-199; pair order feature_detection.hpp:87-90): keypoints, descriptors, both match
lists, DMatch.distance bits, maps_of_indices, the row-band keep flags and the index quadruples.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

H, W = 376, 1240


def _oracle_frame(O, semi, desc, f, K, mode, nthreads=16):
    """Reference result of frame f of a stream (semi/desc: CPU float32 arrays of frames f-1 and f, or of f only
    when f is the first frame after a reset): dict like tests/test_gpu_stereo.py::_oracle_stream."""
    cur = O.decode(semi[-1], desc[-1], max_keypoints=K, num_threads=nthreads)
    nl, nr = int(cur["n"][0]), int(cur["n"][1])
    ms, maps = O.match(cur["desc"][0, :nl], cur["desc"][1, :nr], mode=mode, num_threads=nthreads)
    keep = O.stereo_filter(cur["kpts"][0], cur["kpts"][1], ms, 2.0, 0.25)
    if len(semi) == 2:
        prev = O.decode(semi[0], desc[0], max_keypoints=K, num_threads=nthreads)
        pl, pr = int(prev["n"][0]), int(prev["n"][1])
        mt, mapt = O.match(cur["desc"][0, :nl], prev["desc"][0, :pl], mode=mode, num_threads=nthreads)
        _, prev_maps = O.match(prev["desc"][0, :pl], prev["desc"][1, :pr], mode=mode, num_threads=nthreads)
        quads = O.consistency(ms, mapt, keep, prev_maps)
    else:
        mt, mapt = np.zeros(0, O.DMATCH_DTYPE), np.full(nl, -1, np.int32)
        quads = np.zeros((0, 4), np.int32)
    return dict(dec=cur, ms=ms, maps=maps, keep=keep, mt=mt, mapt=mapt, quads=quads)


def _check_frame(S, out, r, f, F, K, tag):
    """out: numpy views of one batch's spvo_stereo_out; f = frame index within the batch."""
    kp = out["kpts"].view(S.KEYPOINT_DTYPE).reshape(2 * F, K)
    mm = out["matches"].view(S.DMATCH_DTYPE).reshape(2 * F, K)
    for eye in range(2):
        n = int(r["dec"]["n"][eye])
        assert out["n_kpts"][2 * f + eye] == n, (tag, eye)
        assert (kp[2 * f + eye, :n] == r["dec"]["kpts"][eye, :n]).all(), (tag, eye, "keypoints")
        if "desc" in out:
            assert (out["desc"][2 * f + eye, :n].view(np.uint32) == r["dec"]["desc"][eye, :n].view(np.uint32)).all(), \
                (tag, eye, "descriptor bits")
    for row, m, mp, what in ((f, r["ms"], r["maps"], "stereo"), (F + f, r["mt"], r["mapt"], "temporal")):
        k = int(out["n_matches"][row])
        assert k == len(m), (tag, what, k, len(m))
        g = mm[row, :k]
        assert (g["queryIdx"] == m["queryIdx"]).all() and (g["trainIdx"] == m["trainIdx"]).all(), (tag, what)
        assert (g["imgIdx"] == 0).all()
        assert (g["distance"].view(np.uint32) == m["distance"].view(np.uint32)).all(), (tag, what, "distance bits")
        assert (out["q2t"][row, : len(mp)] == mp).all(), (tag, what, "q2t")
    assert (out["stereo_keep"][f, : len(r["keep"])].astype(bool) == r["keep"]).all(), (tag, "keep")
    nq = int(out["n_quads"][f])
    assert nq == len(r["quads"]), (tag, nq, len(r["quads"]))
    assert (out["quads"][f, :nq] == r["quads"]).all(), (tag, "quads")


def _slice_cpu(t, lo, hi):
    return t[lo:hi].float().cpu().numpy()


def _run_config(S, O, K, mode, F, frames, f16=False, host=False, seed=5):
    """Two consecutive batches of F frames through ONE handle; `frames` = global frame indices to check."""
    import torch
    import spvo_b200.synth as synth
    dev = torch.device("cuda", 0)
    semi, desc = synth.make_stream(2 * F, H, W, seed=seed, device=dev)
    if f16:
        semi, desc = semi.half(), desc.half()
    fe = S.Frontend(0, 2 * F, H, W, K)
    fe.set_stream(torch.cuda.current_stream().cuda_stream)
    kw = dict(max_keypoints=K, mode=mode, ratio=0.8, stereo_threshold=2.0, min_disparity=0.25, f16=f16)
    outs = []
    for b in range(2):
        if host:
            hs, hd = semi[b * F:(b + 1) * F].cpu().numpy(), desc[b * F:(b + 1) * F].cpu().numpy()
            out = {k: v.numpy() for k, v in fe.alloc_stereo_out(F, K, device="cpu").items()}
            fe.stereo_batch(hs, hd, F, H, W, out, **kw)
        else:
            dout = fe.alloc_stereo_out(F, K, device=dev)
            fe.stereo_batch_device(semi[b * F:(b + 1) * F], desc[b * F:(b + 1) * F], F, H, W, dout, **kw)
            torch.cuda.synchronize()
            out = {k: v.cpu().numpy() for k, v in dout.items()}
        outs.append(out)
    counters = fe.debug_counters()
    fe.close()
    total_m = 0
    for g in frames:
        lo = max(g - 1, 0)
        r = _oracle_frame(O, _slice_cpu(semi, lo, g + 1), _slice_cpu(desc, lo, g + 1), g, K, mode)
        _check_frame(S, outs[g // F], r, g % F, F, K, f"frame {g}")
        total_m += len(r["ms"]) + len(r["mt"])
    assert total_m > 50 * len(frames), "the synthetic stream must produce real matches"
    return outs, counters


def test_config2_f148_nn_crosscheck_device(spvo, oracle):
    """bench.py's own call: spvo_stereo_batch_device, 148 pairs, K = 1000, NN + cross-check, two batches (carry)."""
    F = 148
    frames = [0, 1, 36, 37, 74, 111, 147, 148, 149, 295]  # first, chunk edges of the host path, last, batch 2 start
    outs, _ = _run_config(spvo, oracle, 1000, 1, F, frames)
    assert (outs[0]["n_kpts"] == 1000).all()


def test_config2_f148_host_chunked(spvo, oracle):
    """The e2e call of bench.py: spvo_stereo_batch (host buffers, 4 chunks with overlapped H2D) at F = 148."""
    F = 148
    frames = [0, 36, 37, 73, 74, 110, 111, 147, 148, 185, 295]  # every quarter-chunk boundary, both sides
    _run_config(spvo, oracle, 1000, 1, F, frames, host=True)


def test_config3_k2048_knn_ratio(spvo, oracle):
    """BASELINE config 3 through the stereo pipeline at KITTI size: K = 2048, kNN-2 + 0.8 ratio test."""
    F = 8
    _run_config(spvo, oracle, 2048, 2, F, list(range(2 * F)))


def test_config2_f16_entry_point(spvo, oracle):
    """spvo_stereo_batch_device_f16 at the bench shape: identical to the oracle on the widened tensors."""
    F = 8
    _run_config(spvo, oracle, 1000, 1, F, list(range(2 * F)), f16=True)


def test_large_image_gather_path_feeds_the_tensor_matcher(spvo, oracle):
    """Planes larger than shared memory (1280x1280: 25 600 cells) take the gather form of descriptor sampling, which
    does not write the tensor matcher's operand slots: the pipeline must convert them itself (k_tc_prep) instead of
    matching on unwritten slots.  Both algorithms must agree with the oracle."""
    import spvo_b200.synth as synth
    S, O = spvo, oracle
    Hh, Ww, K, F = 1280, 1280, 500, 2
    semi, desc = synth.make_stream(2 * F, Hh, Ww, seed=2, device="cpu")
    semi, desc = semi.numpy(), desc.numpy()
    for alg in (S.MATCHER_TENSOR, S.MATCHER_EXACT_FP32):
        fe = S.Frontend(0, 2 * F, Hh, Ww, K)
        for b in range(2):
            out = {k: v.numpy() for k, v in fe.alloc_stereo_out(F, K, device="cpu").items()}
            fe.stereo_batch(semi[b * F:(b + 1) * F], desc[b * F:(b + 1) * F], F, Hh, Ww, out, max_keypoints=K, mode=1,
                            algorithm=alg)
            for f in range(F):
                g = b * F + f
                lo = max(g - 1, 0)
                r = _oracle_frame(O, semi[lo:g + 1], desc[lo:g + 1], g, K, 1)
                _check_frame(S, out, r, f, F, K, f"alg {alg} frame {g}")
        fe.close()
