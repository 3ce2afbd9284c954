import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def spvo():
    import spvo_b200
    return spvo_b200


def make_inputs(B, H, W, seed=0, sigma=1.0):
    """Seeded synthetic SuperPoint head outputs: semi ~ N(0, sigma^2), desc unit-norm per cell."""
    rng = np.random.default_rng(seed)
    Hc, Wc = H // 8, W // 8
    semi = (rng.standard_normal((B, 65, Hc, Wc)) * sigma).astype(np.float32)
    desc = rng.standard_normal((B, 256, Hc, Wc)).astype(np.float32)
    desc /= np.linalg.norm(desc, axis=1, keepdims=True)
    return semi, desc.astype(np.float32)


def unit_rows(n, seed=0, dim=256):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((n, dim)).astype(np.float32)
    return (a / np.linalg.norm(a, axis=1, keepdims=True)).astype(np.float32)
