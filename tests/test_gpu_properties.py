"""Size-independent properties at BASELINE.json's full batch size (74 KITTI-shaped stereo pairs, the bench
workload), where re-running the oracle on every image would take minutes: ordering, NMS spacing, border,
unit norms, match-list structure, cross-check symmetry, index-map / quadruple consistency, determinism."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

H, W, K, F = 376, 1240, 1000, 74


@pytest.fixture(scope="module")
def batch(spvo):
    import torch
    import spvo_b200.synth as synth
    semi, desc = synth.make_stream(F, H, W, seed=7, device="cuda")
    fe = spvo.Frontend(0, 2 * F, H, W, K)
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    fe.set_stream(st.cuda_stream)
    outs = []
    for rep in range(2):  # the same batch twice from a reset state: results must be byte-identical
        fe.stereo_reset()
        o = fe.alloc_stereo_out(F, K, device="cuda")
        sc = torch.zeros(2 * F, K, device="cuda")
        fe.stereo_batch_device(semi, desc, F, H, W, o, max_keypoints=K, mode=spvo.MATCH_NN_CROSSCHECK)
        fe.decode_device(semi.view(2 * F, 65, H // 8, W // 8), desc.view(2 * F, 256, H // 8, W // 8), 2 * F, H, W,
                         torch.zeros_like(o["kpts"]), torch.zeros_like(o["desc"]), torch.zeros_like(o["n_kpts"]), sc,
                         max_keypoints=K)
        torch.cuda.synchronize()
        r = {k: v.cpu().numpy() for k, v in o.items()}
        r["scores"] = sc.cpu().numpy()
        outs.append(r)
    yield spvo, fe, outs, semi, desc
    fe.close()


def test_determinism(batch):
    _, _, outs, _, _ = batch
    for k in outs[0]:
        assert np.array_equal(outs[0][k].view(np.uint8), outs[1][k].view(np.uint8)), k


def test_keypoint_order_spacing_border(batch):
    spvo, _, outs, _, _ = batch
    o = outs[0]
    kp = o["kpts"].view(spvo.KEYPOINT_DTYPE).reshape(2 * F, K)
    assert (o["n_kpts"] == K).all()     # this workload always has more survivors than K
    for b in range(0, 2 * F, 7):
        n = int(o["n_kpts"][b])
        x, y, s = kp[b, :n]["x"], kp[b, :n]["y"], o["scores"][b, :n]
        assert (np.diff(s) <= 0).all() and (s > np.float32(0.015)).all()          # descending score, strict threshold
        assert (x >= 4).all() and (x + 4 < W).all() and (y >= 4).all() and (y + 4 < H).all()   # border_remove
        assert (kp[b, :n]["size"] == 1).all() and (kp[b, :n]["angle"] == -1).all() and (kp[b, :n]["class_id"] == -1).all()
        # greedy NMS: no two kept points within Chebyshev distance 4 (checked on a pixel grid)
        grid = np.zeros((H, W), np.int32)
        grid[y.astype(int), x.astype(int)] = 1
        assert grid.sum() == n                                                      # distinct pixels
        ii = np.cumsum(np.cumsum(np.pad(grid, 5), 0), 1)                             # 9x9 box sums around every keypoint
        yy, xx = y.astype(int) + 5, x.astype(int) + 5
        box = ii[yy + 4, xx + 4] - ii[yy - 5, xx + 4] - ii[yy + 4, xx - 5] + ii[yy - 5, xx - 5]
        assert (box == 1).all()


def test_descriptors_unit_norm(batch):
    _, _, outs, _, _ = batch
    d = outs[0]["desc"]
    nrm = np.linalg.norm(d.astype(np.float64), axis=2)
    assert np.abs(nrm - 1).max() < 1e-6


def test_match_lists_and_maps(batch):
    spvo, fe, outs, _, _ = batch
    o = outs[0]
    m = o["matches"].view(spvo.DMATCH_DTYPE).reshape(2 * F, K)
    assert o["n_matches"][F] == 0 and (o["n_matches"][:F] > 800).all() and (o["n_matches"][F + 1:] > 800).all()
    for row in range(0, 2 * F, 5):
        n = int(o["n_matches"][row])
        q, t, dist = m[row, :n]["queryIdx"], m[row, :n]["trainIdx"], m[row, :n]["distance"]
        assert (np.diff(q) > 0).all() and (m[row, :n]["imgIdx"] == 0).all()          # ascending queryIdx
        assert len(np.unique(t)) == n                                                # cross-check => injective
        q2t = o["q2t"][row]
        exp = np.full(K, -1, np.int32)
        exp[q] = t
        assert (q2t == exp).all()
        # distance field = OpenCV-order fp32 distance of the matched rows (re-derived in float64)
        f = row if row < F else row - F
        a = o["desc"][2 * f][q]
        bslot = 2 * f + 1 if row < F else 2 * (f - 1)
        b = o["desc"][bslot][t]
        ref = np.sqrt(((a.astype(np.float64) - b.astype(np.float64)) ** 2).sum(1))
        assert np.abs(dist - ref).max() < 1e-6
    # quadruples: consistent with matches, maps and the stereo filter
    for f in range(1, F, 9):
        nq = int(o["n_quads"][f])
        qd = o["quads"][f, :nq]
        assert nq > 500 and (np.diff(qd[:, 0]) > 0).all()
        assert (o["q2t"][f][qd[:, 0]] == qd[:, 1]).all()            # currL -> currR
        assert (o["q2t"][F + f][qd[:, 0]] == qd[:, 2]).all()        # currL -> prevL
        assert (o["q2t"][f - 1][qd[:, 2]] == qd[:, 3]).all()        # prevL -> prevR


def test_cross_check_is_symmetric(batch):
    """Matching t against q must give the mirrored match set (mutual nearest neighbours)."""
    spvo, fe, outs, _, _ = batch
    o = outs[0]
    for f in (0, 33, 73):
        a, b = o["desc"][2 * f], o["desc"][2 * f + 1]
        fwd, _ = fe.match(a, b, mode=spvo.MATCH_NN_CROSSCHECK)
        rev, _ = fe.match(b, a, mode=spvo.MATCH_NN_CROSSCHECK)
        s1 = set(zip(fwd["queryIdx"].tolist(), fwd["trainIdx"].tolist()))
        s2 = set(zip(rev["trainIdx"].tolist(), rev["queryIdx"].tolist()))
        assert s1 == s2
        d1 = dict(zip(zip(fwd["queryIdx"].tolist(), fwd["trainIdx"].tolist()), fwd["distance"].view(np.uint32).tolist()))
        d2 = dict(zip(zip(rev["trainIdx"].tolist(), rev["queryIdx"].tolist()), rev["distance"].view(np.uint32).tolist()))
        assert d1 == d2                                           # D(a,b) is bitwise symmetric
