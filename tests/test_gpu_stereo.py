"""GPU parity of the batched stereo-stream entry points (spvo_stereo_batch[_device]) against the
oracle run frame by frame the way the reference's stereoCallback does (visual_odometry_node.cpp:175-199)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle_stream(O, semi, desc, K, mode, thr, mind):
    """semi [F,2,65,Hc,Wc] -> per-frame oracle results."""
    F = semi.shape[0]
    res, prev = [], None
    for f in range(F):
        d = O.decode(semi[f], desc[f], max_keypoints=K)
        nl, nr = int(d["n"][0]), int(d["n"][1])
        ms, maps = O.match(d["desc"][0, :nl], d["desc"][1, :nr], mode=mode)
        keep = O.stereo_filter(d["kpts"][0], d["kpts"][1], ms, thr, mind)
        if prev is not None:
            mt, mapt = O.match(d["desc"][0, :nl], prev["desc"][0, : int(prev["n"][0])], mode=mode)
        else:
            mt, mapt = np.zeros(0, O.DMATCH_DTYPE), np.full(nl, -1, np.int32)
        quads = (O.consistency(ms, mapt, keep, res[-1]["maps"]) if res else np.zeros((0, 4), np.int32))
        res.append(dict(dec=d, ms=ms, maps=maps, keep=keep, mt=mt, mapt=mapt, quads=quads))
        prev = d
    return res


def _check_batch(S, out, ref, f0, F, K, no_carry=False):
    kp = out["kpts"].view(S.KEYPOINT_DTYPE).reshape(2 * F, K)
    mm = out["matches"].view(S.DMATCH_DTYPE).reshape(2 * F, K)
    for f in range(F):
        r = ref[f0 + f]
        for eye in range(2):
            n = int(r["dec"]["n"][eye])
            assert out["n_kpts"][2 * f + eye] == n
            assert (kp[2 * f + eye, :n] == r["dec"]["kpts"][eye, :n]).all()
            if "desc" in out:
                assert (out["desc"][2 * f + eye, :n] == r["dec"]["desc"][eye, :n]).all()
        for row, m, mp in ((f, r["ms"], r["maps"]), (F + f, r["mt"], r["mapt"])):
            k = int(out["n_matches"][row])
            assert k == len(m), (row, k, len(m))
            g = mm[row, :k]
            assert (g["queryIdx"] == m["queryIdx"]).all() and (g["trainIdx"] == m["trainIdx"]).all()
            assert (g["distance"].view(np.uint32) == m["distance"].view(np.uint32)).all()
            assert (out["q2t"][row, : len(mp)] == mp).all()
        assert (out["stereo_keep"][f, : len(r["keep"])].astype(bool) == r["keep"]).all()
        if "quads" in out and not (f0 + f > 0 and f == 0 and no_carry):
            nq = int(out["n_quads"][f])
            assert nq == len(r["quads"]), (f, nq, len(r["quads"]))
            assert (out["quads"][f, :nq] == r["quads"]).all()


@pytest.mark.parametrize("mode", [1, 2])
def test_stereo_batch_host_and_device(spvo, oracle, mode):
    import torch
    import spvo_b200.synth as synth
    S = spvo
    H, W, K, F, NB = 192, 640, 500, 3, 2
    semi, desc = synth.make_stream(F * NB, H, W, seed=3, device="cpu")
    semi, desc = semi.numpy(), desc.numpy()
    ref = _oracle_stream(oracle, semi, desc, K, mode, 2.0, 0.25)
    assert sum(len(r["ms"]) for r in ref) > 100 and sum(len(r["mt"]) for r in ref) > 100
    kw = dict(max_keypoints=K, mode=mode, stereo_threshold=2.0, min_disparity=0.25)

    # host-pointer form, two consecutive batches (the second one's first temporal match uses the carry)
    fe = S.Frontend(0, 2 * F, H, W, K)
    for b in range(NB):
        out = {k: v.numpy() for k, v in fe.alloc_stereo_out(F, K, device="cpu").items()}
        fe.stereo_batch(semi[b * F:(b + 1) * F], desc[b * F:(b + 1) * F], F, H, W, out, **kw)
        _check_batch(S, out, ref, b * F, F, K)
    # reset forgets the previous frame: batch 1 alone has no temporal match for its first frame
    fe.stereo_reset()
    out = {k: v.numpy() for k, v in fe.alloc_stereo_out(F, K, device="cpu").items()}
    fe.stereo_batch(semi[F:2 * F], desc[F:2 * F], F, H, W, out, **kw)
    assert out["n_matches"][F] == 0 and out["n_matches"][F + 1] == len(ref[F + 1]["mt"])
    assert out["n_quads"][0] == 0 and out["n_quads"][1] == len(ref[F + 1]["quads"]) > 20
    fe.close()

    # device-pointer form on the torch stream
    fe = S.Frontend(0, 2 * F, H, W, K)
    fe.set_stream(torch.cuda.current_stream().cuda_stream)
    ds, dd = torch.from_numpy(semi).cuda(), torch.from_numpy(desc).cuda()
    for b in range(NB):
        dout = fe.alloc_stereo_out(F, K, device="cuda")
        # AUTO picks by problem size (api.cu pick_algorithm); the device form pins the tensor-core matcher
        fe.stereo_batch_device(ds[b * F:(b + 1) * F], dd[b * F:(b + 1) * F], F, H, W, dout, algorithm=S.MATCHER_TENSOR, **kw)
        torch.cuda.synchronize()
        _check_batch(S, {k: v.cpu().numpy() for k, v in dout.items()}, ref, b * F, F, K)
    fe.close()


def test_stereo_batch_algorithm_switch_keeps_the_carry(spvo, oracle):
    """Consecutive batches on different matcher algorithms (exact fp32 <-> tensor) continue one sequence."""
    import spvo_b200.synth as synth
    S = spvo
    H, W, K, F = 192, 640, 500, 2
    semi, desc = synth.make_stream(3 * F, H, W, seed=8, device="cpu")
    semi, desc = semi.numpy(), desc.numpy()
    ref = _oracle_stream(oracle, semi, desc, K, 1, 2.0, 0.25)
    fe = S.Frontend(0, 2 * F, H, W, K)
    for b, alg in enumerate((S.MATCHER_EXACT_FP32, S.MATCHER_TENSOR, S.MATCHER_TENSOR)):
        out = {k: v.numpy() for k, v in fe.alloc_stereo_out(F, K, device="cpu").items()}
        fe.stereo_batch(semi[b * F:(b + 1) * F], desc[b * F:(b + 1) * F], F, H, W, out, max_keypoints=K, mode=1,
                        algorithm=alg)
        _check_batch(S, out, ref, b * F, F, K)
    fe.close()


def test_stereo_batch_host_chunked_path(spvo, oracle):
    """F >= 16: the host-pointer call processes the batch as 4 chunks with overlapped H2D; the caller-visible
    whole-batch layout (stereo rows 0..F-1, temporal rows F..2F-1, quadruples) must be unchanged."""
    import spvo_b200.synth as synth
    S = spvo
    H, W, K, F = 64, 96, 60, 18
    semi, desc = synth.make_stream(2 * F, H, W, seed=21, device="cpu")
    semi, desc = semi.numpy(), desc.numpy()
    ref = _oracle_stream(oracle, semi, desc, K, 1, 2.0, 0.25)
    fe = S.Frontend(0, 2 * F, H, W, K)
    for b in range(2):
        out = {k: v.numpy() for k, v in fe.alloc_stereo_out(F, K, device="cpu").items()}
        fe.stereo_batch(semi[b * F:(b + 1) * F], desc[b * F:(b + 1) * F], F, H, W, out, max_keypoints=K, mode=1)
        _check_batch(S, out, ref, b * F, F, K)
    assert sum(len(r["quads"]) for r in ref) > 50
    fe.close()


@pytest.mark.parametrize("K", [333, 130])
def test_stereo_batch_odd_keypoint_budget(spvo, oracle, K):
    """K that is not a multiple of 4 / 16 / 256: the scratch pitch, the padded matcher slots (cap = K rounded to
    256) and the unaligned segments of the one-launch carry copy, on both matcher algorithms."""
    import spvo_b200.synth as synth
    S = spvo
    H, W, F, NB = 128, 320, 2, 2
    semi, desc = synth.make_stream(F * NB, H, W, seed=21, device="cpu")
    semi, desc = semi.numpy(), desc.numpy()
    ref = _oracle_stream(oracle, semi, desc, K, 1, 2.0, 0.25)
    assert all(int(r["dec"]["n"][0]) == K for r in ref), "the budget must bind so that n == K (odd row counts)"
    for alg in (S.MATCHER_TENSOR, S.MATCHER_EXACT_FP32):
        fe = S.Frontend(0, 2 * F, H, W, K)
        for b in range(NB):
            out = {k: v.numpy() for k, v in fe.alloc_stereo_out(F, K, device="cpu").items()}
            fe.stereo_batch(semi[b * F:(b + 1) * F], desc[b * F:(b + 1) * F], F, H, W, out, max_keypoints=K, mode=1,
                            algorithm=alg)
            _check_batch(S, out, ref, b * F, F, K)
        fe.close()


def test_graph_mode_replays_identically(spvo):
    """spvo_set_graph_mode: one pair per call into FIXED bindings (the reference's real-time loop,
    visual_odometry_node.cpp:150-262).  The captured graphs must reproduce the plain launches bit for bit over a
    sequence that includes a reset, and the signature changes (other buffers) must re-capture, not misfire."""
    import torch
    import spvo_b200.synth as synth
    S = spvo
    H, W, K, NF = 192, 640, 500, 14
    dev = torch.device("cuda", 0)
    semi, desc = synth.make_stream(NF, H, W, seed=9, device=dev)
    kw = dict(max_keypoints=K, mode=1, stereo_threshold=2.0, min_disparity=0.25)

    def run(graph):
        fe = S.Frontend(0, 2, H, W, K)
        fe.set_stream(torch.cuda.current_stream().cuda_stream)
        fe.set_graph_mode(graph)
        bs, bd = torch.empty_like(semi[0]), torch.empty_like(desc[0])      # the engine's fixed output bindings
        bs2, bd2 = torch.empty_like(semi[0]), torch.empty_like(desc[0])    # a second set (signature change)
        out = fe.alloc_stereo_out(1, K, device=dev)
        res = []
        for f in range(NF):
            if f == 9:
                fe.stereo_reset()
            s_, d_ = (bs2, bd2) if f in (5, 6) else (bs, bd)
            s_.copy_(semi[f])
            d_.copy_(desc[f])
            fe.stereo_batch_device(s_, d_, 1, H, W, out, **kw)
            torch.cuda.synchronize()
            res.append({k: v.cpu().numpy().copy() for k, v in out.items()})
        launches = fe.kernel_launches
        fe.close()
        return res, launches

    plain, l0 = run(False)
    graph, l1 = run(True)
    assert l0 == l1, "replays must account for the same kernel launches"
    for f, (a, b) in enumerate(zip(plain, graph)):
        for k in a:
            assert a[k].tobytes() == b[k].tobytes(), (f, k)
    assert plain[3]["n_matches"][1] > 50 and plain[9]["n_matches"][1] == 0  # temporal matches; none right after a reset


def test_graph_mode_batches_over_a_ring_of_inputs(spvo):
    """bench.py's shape in small: batches cycling through a ring of input buffers (several call signatures, both carry
    parities) replayed as library-owned graphs must reproduce the plain-launch sequence bit for bit, including the
    concurrent fallback / distance-fill branch and the per-slot detection history."""
    import torch
    import spvo_b200.synth as synth
    S = spvo
    H, W, K, F, R, NB = 192, 640, 500, 6, 3, 16
    dev = torch.device("cuda", 0)
    semi, desc = synth.make_stream(R * F, H, W, seed=21, device=dev)
    semi = semi.view(R, F, 2, 65, H // 8, W // 8)
    desc = desc.view(R, F, 2, 256, H // 8, W // 8)
    kw = dict(max_keypoints=K, mode=1, stereo_threshold=2.0, min_disparity=0.25, algorithm=S.MATCHER_TENSOR)

    def run(graph):
        fe = S.Frontend(0, 2 * F, H, W, K)
        fe.set_stream(torch.cuda.current_stream().cuda_stream)
        fe.set_graph_mode(graph)
        out = fe.alloc_stereo_out(F, K, device=dev)
        res = []
        for b in range(NB):
            fe.stereo_batch_device(semi[b % R], desc[b % R], F, H, W, out, **kw)
            torch.cuda.synchronize()
            res.append({k: v.cpu().numpy().copy() for k, v in out.items()})
        launches = fe.kernel_launches
        fe.close()
        return res, launches

    plain, l0 = run(False)
    graph, l1 = run(True)
    assert l0 == l1
    for b, (a, g) in enumerate(zip(plain, graph)):
        for k in a:
            assert a[k].tobytes() == g[k].tobytes(), (b, k)
    assert plain[-1]["n_matches"].min() > 50
