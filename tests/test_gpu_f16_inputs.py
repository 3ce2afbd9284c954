"""fp16 network outputs (staging straight from an fp16 inference engine): the *_f16 entry points must
give, bit for bit, what the oracle gives on the same values widened to fp32 (binary16 -> binary32 is exact)."""
import numpy as np
import pytest

from conftest import make_inputs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("H,W,K", [(376, 1240, 1000), (192, 640, 500), (128, 320, 333), (64, 72, 50)])
def test_decode_f16_host_and_device(spvo, oracle, H, W, K):
    import torch
    B = 3
    semi, desc = make_inputs(B, H, W, seed=H + 3 * K)
    s16, d16 = semi.astype(np.float16), desc.astype(np.float16)   # odd plane sizes: every 16-byte phase occurs
    ref = oracle.decode(s16.astype(np.float32), d16.astype(np.float32), max_keypoints=K)
    fe = spvo.Frontend(0, B, H, W, K)
    r = fe.decode(s16, d16, max_keypoints=K)
    for b in range(B):
        n = int(ref["n"][b])
        assert r["n"][b] == n
        assert (r["kpts"][b, :n] == ref["kpts"][b, :n]).all()
        assert (r["desc"][b, :n].view(np.uint32) == ref["desc"][b, :n].view(np.uint32)).all()
        assert (r["scores"][b, :n].view(np.uint32) == ref["scores"][b, :n].view(np.uint32)).all()
    # device form, on the torch stream
    fe.set_stream(torch.cuda.current_stream().cuda_stream)
    ds, dd = torch.from_numpy(s16).cuda(), torch.from_numpy(d16).cuda()
    kp = torch.zeros(B, K, 7, dtype=torch.float32, device="cuda")
    do = torch.zeros(B, K, 256, dtype=torch.float32, device="cuda")
    nn = torch.zeros(B, dtype=torch.int32, device="cuda")
    fe.decode_device(ds, dd, B, H, W, kp, do, nn, None, max_keypoints=K, f16=True)
    torch.cuda.synchronize()
    assert (nn.cpu().numpy() == ref["n"]).all()
    for b in range(B):
        n = int(ref["n"][b])
        assert (do[b, :n].cpu().numpy().view(np.uint32) == ref["desc"][b, :n].view(np.uint32)).all()
    fe.close()


def _same_valid(S, a, b, F, K):
    """Equality of two spvo_stereo_out dicts over their valid entries (rows beyond the counts are unspecified)."""
    for k in ("n_kpts", "n_matches", "n_quads"):
        assert (a[k] == b[k]).all(), k
    ka, kb = a["kpts"].view(S.KEYPOINT_DTYPE).reshape(2 * F, K), b["kpts"].view(S.KEYPOINT_DTYPE).reshape(2 * F, K)
    ma, mb = a["matches"].view(S.DMATCH_DTYPE).reshape(2 * F, K), b["matches"].view(S.DMATCH_DTYPE).reshape(2 * F, K)
    for i in range(2 * F):
        n, m = int(a["n_kpts"][i]), int(a["n_matches"][i])
        assert (ka[i, :n] == kb[i, :n]).all()
        assert (a["desc"][i, :n].view(np.uint32) == b["desc"][i, :n].view(np.uint32)).all()
        assert (ma[i, :m].view(np.uint8) == mb[i, :m].view(np.uint8)).all()
    for f in range(F):
        nl = int(a["n_kpts"][2 * f])
        assert (a["q2t"][f, :nl] == b["q2t"][f, :nl]).all() and (a["q2t"][F + f, :nl] == b["q2t"][F + f, :nl]).all()
        ms = int(a["n_matches"][f])
        assert (a["stereo_keep"][f, :ms] == b["stereo_keep"][f, :ms]).all()
        nq = int(a["n_quads"][f])
        assert (a["quads"][f, :nq] == b["quads"][f, :nq]).all()


def test_stereo_batch_f16_matches_the_widened_fp32_run(spvo):
    """The whole stereo pipeline on fp16 tensors == the same pipeline on the widened fp32 tensors (which the other
    tests tie to the oracle), host and device forms, across two batches (carry)."""
    import torch
    import spvo_b200.synth as synth
    S = spvo
    H, W, K, F, NB = 192, 640, 500, 3, 2
    semi, desc = synth.make_stream(F * NB, H, W, seed=13, device="cpu")
    s16, d16 = semi.numpy().astype(np.float16), desc.numpy().astype(np.float16)
    s32, d32 = s16.astype(np.float32), d16.astype(np.float32)
    fa, fb, fc = (S.Frontend(0, 2 * F, H, W, K) for _ in range(3))
    fc.set_stream(torch.cuda.current_stream().cuda_stream)
    for b in range(NB):
        sl = slice(b * F, (b + 1) * F)
        oa = {k: v.numpy() for k, v in fa.alloc_stereo_out(F, K, device="cpu").items()}
        ob = {k: v.numpy() for k, v in fb.alloc_stereo_out(F, K, device="cpu").items()}
        fa.stereo_batch(s32[sl], d32[sl], F, H, W, oa, max_keypoints=K)
        fb.stereo_batch(s16[sl], d16[sl], F, H, W, ob, max_keypoints=K, f16=True)
        oc = fc.alloc_stereo_out(F, K, device="cuda")
        fc.stereo_batch_device(torch.from_numpy(s16[sl]).cuda(), torch.from_numpy(d16[sl]).cuda(), F, H, W, oc,
                               max_keypoints=K, f16=True)
        torch.cuda.synchronize()
        assert oa["n_matches"].sum() > 500
        for o in (ob, {k: v.cpu().numpy() for k, v in oc.items()}):
            _same_valid(S, oa, o, F, K)
    for f in (fa, fb, fc):
        f.close()


def test_stereo_batch_f16_host_chunked_path(spvo):
    """F >= 16: the host form streams the batch as 4 chunks; with fp16 inputs every chunk offset is in 2-byte
    elements.  Must equal the fp32 host call on the widened tensors."""
    import spvo_b200.synth as synth
    S = spvo
    H, W, K, F = 128, 320, 200, 18
    semi, desc = synth.make_stream(F, H, W, seed=31, device="cpu")
    s16, d16 = semi.numpy().astype(np.float16), desc.numpy().astype(np.float16)
    fa, fb = S.Frontend(0, 2 * F, H, W, K), S.Frontend(0, 2 * F, H, W, K)
    oa = {k: v.numpy() for k, v in fa.alloc_stereo_out(F, K, device="cpu").items()}
    ob = {k: v.numpy() for k, v in fb.alloc_stereo_out(F, K, device="cpu").items()}
    fa.stereo_batch(s16.astype(np.float32), d16.astype(np.float32), F, H, W, oa, max_keypoints=K)
    fb.stereo_batch(s16, d16, F, H, W, ob, max_keypoints=K, f16=True)
    assert oa["n_matches"].sum() > 1000
    _same_valid(S, oa, ob, F, K)
    fa.close()
    fb.close()
