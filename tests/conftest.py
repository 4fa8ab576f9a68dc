import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `pytest -m gpu` on the GPU box)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


_CACHE = {}


def workload(name, **kw):
    """Cached (workload, host operands).  ``b_order=1``: the same set-up with P1 buoyancy."""
    from nupgcm_b200 import workloads as W
    key = (name, tuple(sorted(kw.items())))
    if key not in _CACHE:
        kw = dict(kw)
        b_order = kw.pop("b_order", None)
        w = getattr(W, name)(**kw)
        if b_order is not None:
            w = W.with_b_order(w, b_order)
        _CACHE[key] = (w, W.host_operands(w))
    return _CACHE[key]


@pytest.fixture(scope="session")
def ctx():
    """Library context on cuda:0 (GPU tests only).  Fails loudly if the library is missing."""
    from nupgcm_b200.architectures import GPU
    return GPU(0).ctx
