import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_ok():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_ok():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def svcdata():
    return np.load(os.path.join(ROOT, "tests", "golden", "svcdata.npz"))


@pytest.fixture(scope="session")
def golden():
    import json
    return json.load(open(os.path.join(ROOT, "tests", "golden", "vignette_golden.json")))


@pytest.fixture(scope="session")
def vignette_train(svcdata):
    ids = svcdata["id_train"] - 1
    return svcdata["y"][ids], svcdata["X"][ids], svcdata["locs"][ids]


def synth(n, p, seed, dim=2, q=None):
    """sample_SVCdata-style synthetic inputs: constant density 10 pts / unit^2."""
    rng = np.random.default_rng(seed)
    L = 10.0 * np.sqrt(n / 1000.0)
    locs = rng.uniform(0, L, size=(n, dim))
    X = np.column_stack([np.ones(n), rng.standard_normal((n, p - 1))]) if p > 1 else np.ones((n, 1))
    q = p if q is None else q
    W = X[:, :q] if q <= p else np.column_stack([X, rng.standard_normal((n, q - p))])
    y = X @ np.arange(1, p + 1) + rng.standard_normal(n) * 1.5
    return locs, X, W, y


THETA10 = [1.3, 0.7, 2.1, 0.4, 0.9, 1.1, 1.7, 0.3, 2.5, 0.2]


def theta_for(q, tau2=0.35):
    return np.array(THETA10[:2 * q] + [tau2])
