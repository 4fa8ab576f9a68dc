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
    """Cached (workload, host operands)."""
    from nupgcm_b200 import workloads as W
    key = (name, tuple(sorted(kw.items())))
    if key not in _CACHE:
        w = getattr(W, name)(**kw)
        _CACHE[key] = (w, W.host_operands(w))
    return _CACHE[key]


@pytest.fixture(scope="session")
def ctx():
    """Library context on cuda:0 (GPU tests only).  Fails loudly if the library is missing."""
    from nupgcm_b200.architectures import GPU
    return GPU(0).ctx
