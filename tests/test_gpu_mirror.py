"""The host-side mirrors of the reference interface (C++ header-only class and its Python twin)
driven exactly like the reference's stereoCallback (visual_odometry_node.cpp:175-199), checked
against the oracle frame by frame."""
import os
import struct
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "test_frontend_mirror")


def _oracle_frames(O, semi, desc, K, mode):
    res, prev = [], None
    for f in range(semi.shape[0]):
        d = O.decode(semi[f], desc[f], max_keypoints=K)
        nl, nr = int(d["n"][0]), int(d["n"][1])
        ms, maps = O.match(d["desc"][0, :nl], d["desc"][1, :nr], mode=mode)
        mt, mapt = (O.match(d["desc"][0, :nl], prev["desc"][0, : int(prev["n"][0])], mode=mode) if prev is not None
                    else (np.zeros(0, O.DMATCH_DTYPE), np.full(nl, -1, np.int32)))
        res.append((d, ms, maps, mt, mapt))
        prev = d
    return res


@pytest.mark.parametrize("selector,cross,batch", [("NN", 1, 2), ("KNN", 0, 1), ("NN", 0, 2)])
def test_cpp_mirror_class(oracle, spvo, tmp_path, selector, cross, batch):
    import spvo_b200.synth as synth
    if not os.path.exists(BIN):
        import __graft_entry__ as g
        g.build()
    H, W, K, F = 192, 640, 400, 3
    semi, desc = synth.make_stream(F, H, W, seed=6, device="cpu")
    semi, desc = semi.numpy(), desc.numpy()
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as f:
        f.write(struct.pack("5i", F, H, W, K, batch))
        for i in range(F):
            f.write(semi[i].tobytes())
            f.write(desc[i].tobytes())
    subprocess.run([BIN, str(fin), str(fout), selector, str(cross)], check=True, timeout=300)
    mode = 2 if selector == "KNN" else (1 if cross else 0)
    ref = _oracle_frames(oracle, semi, desc, K, mode)
    buf = open(fout, "rb").read()
    off = 0

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(buf, dtype, count, off)
        off += a.nbytes
        return a

    for f in range(F):
        d, ms, maps, mt, mapt = ref[f]
        nl, nr = take(np.int32, 2)
        assert (nl, nr) == (d["n"][0], d["n"][1])
        kl, kr = take(spvo.KEYPOINT_DTYPE, nl), take(spvo.KEYPOINT_DTYPE, nr)
        assert (kl == d["kpts"][0, :nl]).all() and (kr == d["kpts"][1, :nr]).all()
        dl, dr = take(np.float32, nl * 256), take(np.float32, nr * 256)
        assert (dl.reshape(nl, 256) == d["desc"][0, :nl]).all() and (dr.reshape(nr, 256) == d["desc"][1, :nr]).all()
        for m, mp in ((ms, maps), (mt, mapt)):
            n = int(take(np.int32, 1)[0])
            g = take(spvo.DMATCH_DTYPE, n)
            gm = take(np.int32, nl)
            assert n == len(m) and (g["queryIdx"] == m["queryIdx"]).all() and (g["trainIdx"] == m["trainIdx"]).all()
            assert (g["distance"].view(np.uint32) == m["distance"].view(np.uint32)).all()
            assert (gm == mp).all()
    assert off == len(buf)


def test_python_mirror_class(oracle, spvo):
    import spvo_b200.synth as synth
    H, W, K, F = 120, 392, 300, 3   # the reference default constructor's size (hpp:255-265)
    semi, desc = synth.make_stream(F, H, W, seed=2, device="cpu")
    semi, desc = semi.numpy(), desc.numpy()
    fe = spvo.SuperPointFeatureFrontEnd("NN", True, 2, H, W, max_keypoints=K)
    ref = _oracle_frames(oracle, semi, desc, K, 1)
    for f in range(F):
        fe.output_det_data_[...] = semi[f]
        fe.output_desc_data_[...] = desc[f]
        fe.postprocessDetectionAndDescription()
        fe.matchDescriptors(spvo.CURR_LEFT_CURR_RIGHT)
        if f > 0:
            fe.matchDescriptors(spvo.CURR_LEFT_PREV_LEFT)
        d, ms, maps, mt, mapt = ref[f]
        assert (fe.keypoints_dq[spvo.CURR_LEFT] == d["kpts"][0, : d["n"][0]]).all()
        assert (fe.descriptors_dq[spvo.CURR_RIGHT] == d["desc"][1, : d["n"][1]]).all()
        assert (fe.cv_DMatches_list[spvo.CURR_LEFT_CURR_RIGHT] == ms).all()
        assert (fe.maps_of_indices[spvo.CURR_LEFT_CURR_RIGHT] == maps).all()
        if f > 0:
            assert (fe.cv_DMatches_list[spvo.CURR_LEFT_PREV_LEFT] == mt).all()
            assert (fe.maps_of_indices[spvo.PREV_LEFT_PREV_RIGHT] == ref[f - 1][2]).all()   # BASE:475-481
        assert len(fe.keypoints_dq) <= 4


def test_cpp_mirror_preprocess_image(oracle, spvo, tmp_path):
    """spvo::SuperPointFeatureFrontEnd::preprocessImage (NN:139-161) through the C++ mirror vs the oracle."""
    if not os.path.exists(BIN):
        import __graft_entry__ as g
        g.build()
    rows, cols, H, W = 370, 1226, 120, 392
    img = np.random.default_rng(9).integers(0, 256, (rows, cols), dtype=np.uint8)
    P = np.array([707.09, 0, 601.89, -379.8, 0, 707.09, 183.11, 0, 0, 0, 1, 0], np.float32)
    fin, fout = tmp_path / "pin.bin", tmp_path / "pout.bin"
    with open(fin, "wb") as f:
        f.write(struct.pack("4i", rows, cols, H, W))
        f.write(P.tobytes())
        f.write(img.tobytes())
    subprocess.run([BIN, "--preprocess", str(fin), str(fout)], check=True, timeout=120)
    buf = open(fout, "rb").read()
    inp = np.frombuffer(buf, np.float32, H * W).reshape(H, W)
    rs = np.frombuffer(buf, np.uint8, H * W, H * W * 4).reshape(H, W)
    Pp = np.frombuffer(buf, np.float32, 12, H * W * 5).reshape(3, 4)
    oi, ors, oP = oracle.preprocess(img, H, W, P)
    assert (rs == ors).all() and (inp.view(np.uint32) == oi.view(np.uint32)).all()
    assert (Pp.view(np.uint32) == oP.view(np.uint32)).all()
