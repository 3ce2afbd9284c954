"""k_softmax_heat divides by ONE correctly rounded reciprocal per cell (Markstein's correction sequence) instead of 64
IEEE divisions.  The quotient must equal the reference's `det / (sum + 1e-5)` true division (NN:280-284) bit for bit:
spvo_debug_div_check compares the two forms on the device for arbitrary operand bit patterns."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check(fe, a, b):
    import torch
    ta = torch.from_numpy(np.ascontiguousarray(a, np.float32).view(np.int32)).cuda()
    tb = torch.from_numpy(np.ascontiguousarray(b, np.float32).view(np.int32)).cuda()
    return fe.div_check(ta, tb)


def test_shared_reciprocal_division_is_ieee_exact(spvo):
    import torch
    fe = spvo.Frontend(0, 1, 64, 64, 16)
    rng = np.random.default_rng(0)
    n = 1 << 24
    # softmax-shaped operands: a = exp(x) of one channel, b = the 65-term sum (+1e-5)
    x = rng.normal(0, 3, n).astype(np.float32)
    a = np.exp(x)
    b = (a + np.exp(rng.normal(0, 3, n).astype(np.float32)) * rng.integers(1, 65, n)).astype(np.float32) + np.float32(1e-5)
    assert _check(fe, a, b) == 0
    # arbitrary positive pairs over the whole guarded range (random bit patterns)
    for seed in range(4):
        g = np.random.default_rng(100 + seed)
        ea, eb = g.integers(127 - 60, 127 + 60, n), g.integers(127 - 59, 127 + 59, n)
        abits = (ea.astype(np.uint32) << 23) | g.integers(0, 1 << 23, n).astype(np.uint32)
        bbits = (eb.astype(np.uint32) << 23) | g.integers(0, 1 << 23, n).astype(np.uint32)
        assert _check(fe, abits.view(np.float32), bbits.view(np.float32)) == 0
    # adversarial significands: all ones, powers of two, one-off patterns, against random numerators
    g = np.random.default_rng(7)
    special = np.array([0x000000, 0x000001, 0x7FFFFF, 0x7FFFFE, 0x400000, 0x3FFFFF, 0x400001, 0x555555, 0x2AAAAA,
                        0x7FF000, 0x000FFF, 0x100000, 0x600000], np.uint32)
    bb = ((np.uint32(127) + g.integers(-20, 20, (n // 16)).astype(np.uint32)) << 23).astype(np.uint32) | \
        special[g.integers(0, len(special), n // 16)]
    for trial in range(4):
        aa = ((np.uint32(127) + g.integers(-30, 30, n // 16).astype(np.uint32)) << 23).astype(np.uint32) | \
            (special[g.integers(0, len(special), n // 16)] if trial % 2 else g.integers(0, 1 << 23, n // 16).astype(np.uint32))
        assert _check(fe, aa.view(np.float32), bb.view(np.float32)) == 0
    # exact quotients and exact halfway cases: a = q * b for short q, b
    q = g.integers(1, 1 << 12, n // 16).astype(np.float32)
    bs = g.integers(1, 1 << 12, n // 16).astype(np.float32)
    assert _check(fe, q * bs, bs) == 0
    assert _check(fe, (q * bs + np.float32(0.5)).astype(np.float32), bs) == 0
    fe.close()
