"""Streaming form of the persistent solvers (matrix too large for shared memory): parity on a
once-refined bowl3D mesh (N ~ 1.3e5, ~7e6 non-zeros), where the SpMV runs through the TMA chunk
pipeline instead of the SM-resident path."""
import os

import numpy as np
import pytest

from conftest import workload
from nupgcm_b200 import lib
from nupgcm_b200 import workloads as W
from oracle import krylov

pytestmark = pytest.mark.gpu

_CACHE = {}


def refined_ops():
    if "ops" not in _CACHE:
        w = W.bowl_example(mesh=W.refined_bowl(1, h0=0.1))
        _CACHE["ops"] = W.host_operands(w)
    return _CACHE["ops"]


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("orth", [lib.ORTH_MGS, lib.ORTH_CGS2, lib.ORTH_CGS2_FUSED])
@pytest.mark.parametrize("tma", ["1", "0"])
def test_streaming_gmres_matches_oracle(ctx, orth, tma):
    ops = refined_ops()
    A = ops["A"]
    assert A.shape[0] > 100000
    y = ops["B"] @ ops["b_init"] + ops["b0"]
    M = np.full(y.size, ops["pscale"])
    itmax = 45                                    # two full restart cycles and a partial one
    xo, so = krylov.gmres(A, y, x0=np.zeros(y.size), M=M, atol=0.0, rtol=1e-30, memory=20, itmax=itmax,
                          orth={lib.ORTH_MGS: "mgs", lib.ORTH_CGS2: "cgs2", lib.ORTH_CGS2_FUSED: "cgs2f"}[orth])
    os.environ["NUPGCM_STREAM_TMA"] = tma
    try:
        x = ctx.vector(y.size)
        st, hist = lib.gmres_solve(ctx.csr(A, drop_zeros=True), ctx.vector(y), x, pscale=ops["pscale"],
                                   atol=0.0, rtol=1e-30, itmax=itmax, memory=20, orth=orth, history=64)
    finally:
        os.environ.pop("NUPGCM_STREAM_TMA")
    assert st.niter == itmax
    assert np.allclose(hist, so.residuals, rtol=1e-9)
    assert rel(x.download(), xo) < 1e-9


def test_streaming_cg_matches_oracle(ctx):
    ops = refined_ops()
    A = (ops["M"] + 1e-3 * (ops["Kh"] + ops["Kv"])).tocsr()
    b = np.random.default_rng(0).uniform(-1, 1, A.shape[0])
    dinv = 1.0 / A.diagonal()
    xo, so = krylov.cg(A, b, x0=np.zeros(b.size), M=dinv, atol=1e-6, rtol=1e-6)
    # force the streaming form on this small matrix by disabling the resident one
    os.environ["NUPGCM_RESIDENT"] = "0"
    try:
        x = ctx.vector(b.size)
        st, hist = lib.cg_solve(ctx.csr(A), ctx.vector(b), x, dinv=ctx.vector(dinv), atol=1e-6, rtol=1e-6, history=512)
    finally:
        os.environ.pop("NUPGCM_RESIDENT")
    assert st.solved and abs(st.niter - so.niter) <= 1
    assert rel(x.download(), xo) < 1e-6
