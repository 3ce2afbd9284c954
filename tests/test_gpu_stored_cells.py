"""k_softmax_heat stores the heat values only of the cells it PREDICTS k_detect will open (per image slot: second
largest value >= 0.9 x the lowest score bound k_detect needed on the previous call); every other cell k_detect needs
is recomputed from the logits.  The prediction may only ever change the time taken: these tests drive one handle
through call sequences whose history is absent, right, too high and too low and compare every call with the oracle.
Reference arithmetic: feature_detection_neural_network.cpp:266-330.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _same(r, o, b_r, b_o):
    n = int(o["n"][b_o])
    assert int(r["n"][b_r]) == n
    assert (r["kpts"][b_r, :n] == o["kpts"][b_o, :n]).all()
    assert (r["scores"][b_r, :n].view(np.uint32) == o["scores"][b_o, :n].view(np.uint32)).all()


@pytest.mark.parametrize("f16", [False, True])
def test_history_never_changes_results(spvo, oracle, f16):
    S, O = spvo, oracle
    import spvo_b200.synth as synth_mod

    H, W, K = 376, 1240, 1000
    real = np.load(os.path.join(GOLD, "realistic_kitti_1240x376.npz"))["semi"][:2]
    synth, _ = synth_mod.make_stream(1, H, W, seed=5)
    synth = synth.reshape(2, 65, H // 8, W // 8).numpy()
    flat = np.full_like(synth, 0.25)  # every pixel 1/65 > conf: one giant tie, radix-select path over recomputed cells
    flat[:, :, :, :4] += np.random.default_rng(0).normal(size=flat[:, :, :, :4].shape).astype(np.float32)
    if f16:
        real, synth, flat = (a.astype(np.float16) for a in (real, synth, flat))
    ref = {name: O.decode(a.astype(np.float32), None, max_keypoints=K, num_threads=8)
           for name, a in (("real", real), ("synth", synth), ("flat", flat))}
    data = dict(real=real, synth=synth, flat=flat)
    fe = S.Frontend(0, 2, H, W, K)
    # no history -> right history -> history from a sparse image (threshold far too low for the dense one: everything
    # is stored) -> history from the dense image (threshold too high for the sparse one: everything is recomputed)
    for name in ("synth", "synth", "real", "real", "synth", "flat", "real", "flat", "synth"):
        a = data[name]
        r = fe.decode(a if f16 else a.astype(np.float32), None, max_keypoints=K)
        for b in range(2):
            _same(r, ref[name], b, b)
    # slots are independent: swap the two images of a batch between calls
    mixed = np.stack([real[0], synth[1]])
    omix = O.decode(mixed.astype(np.float32), None, max_keypoints=K, num_threads=8)
    for _ in range(2):
        r = fe.decode(mixed if f16 else mixed.astype(np.float32), None, max_keypoints=K)
        for b in range(2):
            _same(r, omix, b, b)
    swapped = mixed[::-1].copy()
    r = fe.decode(swapped if f16 else swapped.astype(np.float32), None, max_keypoints=K)
    _same(r, omix, 0, 1)
    _same(r, omix, 1, 0)
    fe.close()


def test_other_thresholds_after_history(spvo, oracle):
    """A handle whose history comes from conf 0.015 / K 1000 is then asked for other parameters."""
    S, O = spvo, oracle
    H, W = 240, 784
    semi = np.load(os.path.join(GOLD, "realistic_kitti_784x240.npz"))["semi"][:2].astype(np.float32)
    fe = S.Frontend(0, 2, H, W, 2048)
    fe.decode(semi, None, max_keypoints=1000)
    for cfg in (dict(max_keypoints=2048, conf_thresh=0.001), dict(max_keypoints=100, conf_thresh=0.2),
                dict(max_keypoints=1000, conf_thresh=0.015, dist_thresh=0), dict(max_keypoints=2048, conf_thresh=0.0005)):
        r = fe.decode(semi, None, **cfg)
        o = O.decode(semi, None, num_threads=8, **cfg)
        for b in range(2):
            _same(r, o, b, b)
    fe.close()


@pytest.mark.parametrize("f16", [False, True])
def test_out_of_range_and_non_finite_logits(spvo, oracle, f16):
    """Logits outside exp's clamp range (|x| > 88.376), huge, infinite and NaN: the packed fast path of k_softmax_heat
    must hand such cells to the exact scalar form (oracle_exp: clamp, Cephes polynomial, max with the argument)."""
    S, O = spvo, oracle
    import spvo_b200.synth as synth_mod

    H, W, K = 240, 784, 600
    semi, _ = synth_mod.make_stream(1, H, W, seed=11)
    semi = semi.reshape(2, 65, H // 8, W // 8).numpy().copy()
    rng = np.random.default_rng(3)
    specials = [88.0, 88.5, -88.0, -88.5, 89.0, -90.0, 100.0, -100.0, 127.0, -127.5, 1e4, -1e4, 6e4, -6e4]
    if not f16:
        specials += [1e30, -1e30, 3e38, -3e38]
    specials += [np.inf, -np.inf, np.nan]
    flat = semi.reshape(-1)
    idx = rng.choice(flat.size, size=4000, replace=False)
    flat[idx] = rng.choice(np.array(specials, np.float32), size=idx.size)
    # a few cells entirely made of large equal logits (sum overflows / saturates) and of very negative ones
    semi[0, :, 3, 5] = 88.0
    semi[0, :, 4, 7] = -88.3
    semi[1, :, 6, 9] = 80.0
    semi[1, :64, 8, 11] = -100.0
    a = semi.astype(np.float16) if f16 else semi
    with np.errstate(all="ignore"):
        o = O.decode(a.astype(np.float32), None, max_keypoints=K, num_threads=8)
    assert (o["n"] > 100).all()
    fe = S.Frontend(0, 2, H, W, K)
    for _ in range(2):  # second call: stored-cell path
        r = fe.decode(a if f16 else a.astype(np.float32), None, max_keypoints=K)
        for b in range(2):
            _same(r, o, b, b)
    fe.close()
