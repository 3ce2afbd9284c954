"""GPU parity: CUDA matcher (through the C ABI) vs the CPU oracle (itself pinned to cv2.BFMatcher).
Bar: (queryIdx, trainIdx) pairs, order, and the fp32 distance bits all exact."""
import numpy as np
import pytest

from conftest import unit_rows

pytestmark = pytest.mark.gpu

MODES = [0, 1, 2]
ALGS = [1, 2]  # SPVO_MATCHER_EXACT_FP32, SPVO_MATCHER_TENSOR (tcgen05)


def _noisy_copy(a, seed, noise=0.05, perm=True):
    rng = np.random.default_rng(seed)
    b = a + noise * rng.standard_normal(a.shape).astype(np.float32)
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    if perm:
        b = b[rng.permutation(len(b))]
    return b.astype(np.float32)


def _check(fe, O, q, t, mode, algorithm=0):
    gm, gmap = fe.match(q, t, mode=mode, ratio=0.8, algorithm=algorithm)
    om, omap = O.match(q, t, mode=mode, ratio=0.8)
    assert len(gm) == len(om), (len(gm), len(om))
    assert (gm["queryIdx"] == om["queryIdx"]).all() and (gm["trainIdx"] == om["trainIdx"]).all()
    assert (gm["imgIdx"] == 0).all()
    assert (gm["distance"].view(np.uint32) == om["distance"].view(np.uint32)).all()
    assert (gmap == omap).all()
    return gm


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("N,M", [(1000, 1000), (777, 1023), (33, 2048), (1, 1), (2, 1), (500, 3)])
@pytest.mark.parametrize("alg", ALGS)
def test_match_parity_structured(spvo, oracle, mode, N, M, alg):
    fe = spvo.Frontend(0, 1, 64, 64, 16)
    base = unit_rows(max(N, M), seed=N * 7 + M)
    q = base[:N]
    t = _noisy_copy(base, seed=3)[:M]
    gm = _check(fe, oracle, q, t, mode, alg)
    if mode == 2 and N >= 500 and M >= 500:
        assert len(gm) > 100  # ratio test is non-degenerate on structured data
    fe.close()


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("alg", ALGS)
def test_match_parity_random(spvo, oracle, mode, alg):
    fe = spvo.Frontend(0, 1, 64, 64, 16)
    _check(fe, oracle, unit_rows(600, 1), unit_rows(640, 2), mode, alg)
    fe.close()


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("alg", ALGS)
def test_match_ties_lowest_index_wins(spvo, oracle, mode, alg):
    fe = spvo.Frontend(0, 1, 64, 64, 16)
    q, t = unit_rows(200, 11), unit_rows(220, 12)
    t[50] = t[7]; t[120] = t[7]; t[121] = t[7]     # duplicate train rows
    q[30] = q[4]; q[31] = q[4]                     # duplicate query rows
    t[7] = q[4]                                    # an exact zero distance, shared
    _check(fe, oracle, q, t, mode, alg)
    fe.close()


def test_match_empty_inputs(spvo):
    fe = spvo.Frontend(0, 1, 64, 64, 16)
    for mode in MODES:
        m, q2t = fe.match(np.zeros((0, 256), np.float32), unit_rows(5), mode=mode)
        assert len(m) == 0 and len(q2t) == 0
        m, q2t = fe.match(unit_rows(5), np.zeros((0, 256), np.float32), mode=mode)
        assert len(m) == 0 and (q2t == -1).all()
    m, q2t = fe.match(unit_rows(5), unit_rows(1), mode=2)   # kNN with one train row: defined as no match
    assert len(m) == 0 and (q2t == -1).all()
    fe.close()


@pytest.mark.parametrize("alg", ALGS)
def test_match_batch_device(spvo, oracle, alg):
    """Batched device API: slots as written by decode, per-slot row counts read on the device."""
    import torch
    fe = spvo.Frontend(0, 1, 64, 64, 16)
    S, stride = 6, 300
    rows = [300, 250, 0, 1, 299, 128]
    descs = np.zeros((S, stride, 256), np.float32)
    base = unit_rows(stride, 5)
    for s in range(S):
        descs[s, : rows[s]] = _noisy_copy(base, seed=s, perm=(s % 2 == 1))[: rows[s]]
    d = torch.from_numpy(descs).cuda()
    n_rows = torch.tensor(rows, dtype=torch.int32, device="cuda")
    qs = torch.tensor([0, 0, 1, 2, 3, 4, 5], dtype=torch.int32, device="cuda")
    ts = torch.tensor([1, 4, 0, 1, 0, 5, 2], dtype=torch.int32, device="cuda")
    P = len(qs)
    fe.set_stream(torch.cuda.current_stream().cuda_stream)
    for mode in MODES:
        out = torch.zeros(P, stride, 4, dtype=torch.int32, device="cuda")
        nm = torch.zeros(P, dtype=torch.int32, device="cuda")
        q2t = torch.zeros(P, stride, dtype=torch.int32, device="cuda")
        fe.match_batch_device(d, n_rows, stride, qs, ts, P, stride, out, nm, q2t, mode=mode, algorithm=alg)
        torch.cuda.synchronize()
        out_h = out.cpu().numpy().view(spvo.DMATCH_DTYPE).reshape(P, stride)
        for p in range(P):
            a, b = int(qs[p]), int(ts[p])
            om, omap = oracle.match(descs[a, : rows[a]], descs[b, : rows[b]], mode=mode)
            k = int(nm[p])
            assert k == len(om)
            g = out_h[p, :k]
            assert (g["queryIdx"] == om["queryIdx"]).all() and (g["trainIdx"] == om["trainIdx"]).all()
            assert (g["distance"].view(np.uint32) == om["distance"].view(np.uint32)).all()
            assert (q2t[p, : rows[a]].cpu().numpy() == omap).all()
    fe.close()


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("N,M", [(2048, 2048), (4096, 3000), (129, 257)])
def test_match_tensor_large(spvo, oracle, mode, N, M):
    """Config 5 sizes through the tcgen05 path; the oracle is the checker."""
    fe = spvo.Frontend(0, 1, 64, 64, 16)
    base = unit_rows(max(N, M), seed=N + M)
    _check(fe, oracle, base[:N], _noisy_copy(base, seed=8)[:M], mode, 2)
    fe.close()


def test_match_tensor_random_uses_fallback_and_stays_exact(spvo, oracle):
    """Unstructured descriptors: many rows cannot be proved from the bf16 shortlist and take the
    exact full-row fallback; results must not change."""
    fe = spvo.Frontend(0, 1, 64, 64, 16)
    q, t = unit_rows(1500, 21), unit_rows(1400, 22)
    for mode in MODES:
        _check(fe, oracle, q, t, mode, 2)
    fb = int(fe.debug_counters()[1])
    print("fallback rows:", fb)
    assert fb > 50, "the test is meant to drive rows (forward and reverse, pending while k_tc_fill_dist runs) through k_tc_fallback"
    fe.close()


def test_match_tensor_max_size_8192(spvo, oracle):
    """Largest size of the packed shortlist index (13 bits): N = M = 8192, cross-check."""
    fe = spvo.Frontend(0, 1, 64, 64, 16)
    base = unit_rows(8192, seed=99)
    q, t = base, _noisy_copy(base, seed=5)
    gm, gmap = fe.match(q, t, mode=1, algorithm=2)
    om, omap = oracle.match(q, t, mode=1, num_threads=16)
    assert len(gm) == len(om) > 7000 and (gm["queryIdx"] == om["queryIdx"]).all() and (gm["trainIdx"] == om["trainIdx"]).all()
    assert (gm["distance"].view(np.uint32) == om["distance"].view(np.uint32)).all() and (gmap == omap).all()
    fe.close()


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("scale", [1e-3, 37.0, 1e4])
def test_match_tensor_arbitrary_descriptor_scale(spvo, oracle, mode, scale):
    """Generic CV_32F descriptors are not unit-norm: the bf16 bound, the key offset and the key quantisation all
    scale with the operands' largest norms.  Mixed row norms (x0.25 .. x4), a few all-zero rows, tensor path."""
    rng = np.random.default_rng(int(scale * 10) % 1000 + mode)
    N, M = 300, 520
    base = unit_rows(M, seed=77)
    q = (base[:N] + 0.05 * rng.standard_normal((N, 256)).astype(np.float32)).astype(np.float32)
    t = base[rng.permutation(M)].copy()
    q *= (np.float32(scale) * rng.choice(np.array([0.25, 1.0, 4.0], np.float32), size=(N, 1))).astype(np.float32)
    t *= (np.float32(scale) * rng.choice(np.array([0.25, 1.0, 4.0], np.float32), size=(M, 1))).astype(np.float32)
    q[5] = 0.0
    t[9] = 0.0
    t[11] = 0.0
    fe = spvo.Frontend(0, 1, 64, 64, 16)
    _check(fe, oracle, q, t, mode, 2)
    fe.close()
