"""Streaming form of the persistent solvers (matrix too large for shared memory): parity on a
once-refined bowl3D mesh (N ~ 1.3e5, ~7e6 non-zeros), where the SpMV runs through the tiled TMA
streams (per-warp rings + staged footprints) instead of the SM-resident path; and the same form
forced onto the shipped meshes with small footprint caps (many tiles per CTA)."""
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


@pytest.mark.parametrize("orth,name,grid", [(lib.ORTH_MGS, "mgs", "4"), (lib.ORTH_MGS, "mgs", "148"),
                                            (lib.ORTH_CGS2_FUSED, "cgs2f", "24"), (lib.ORTH_CGS2_FUSED, "cgs2f", "148")])
def test_forced_streaming_many_tiles_per_cta(ctx, orth, name, grid):
    """h=0.1 3-D inversion matrix through the streaming form on few CTAs: a dozen tiles per CTA, so the
    footprint arena is gone round several times and every ring slot is recycled many times per SpMV."""
    _, ops = workload("bowl_mixing")
    A = ops["A"]
    rng = np.random.default_rng(5)
    y = rng.uniform(-1, 1, A.shape[0])
    x0 = rng.uniform(-1, 1, A.shape[0])
    M = np.full(y.size, ops["pscale"])
    xo, so = krylov.gmres(A, y, x0=x0, M=M, atol=0.0, rtol=1e-30, memory=20, itmax=45, orth=name)
    os.environ["NUPGCM_RESIDENT"] = "0"
    os.environ["NUPGCM_GRID"] = grid
    try:
        dA = ctx.csr(A, drop_zeros=True)
        runs = []
        for _ in range(2):
            x = ctx.vector(x0)
            st, hist = lib.gmres_solve(dA, ctx.vector(y), x, pscale=ops["pscale"], atol=0.0, rtol=1e-30,
                                       itmax=45, memory=20, orth=orth, history=64)
            runs.append((hist.copy(), x.download()))
    finally:
        os.environ.pop("NUPGCM_RESIDENT")
        os.environ.pop("NUPGCM_GRID")
    assert st.niter == 45
    assert np.allclose(runs[0][0], so.residuals, rtol=1e-9)
    assert rel(runs[0][1], xo) < 1e-9
    assert np.array_equal(runs[0][1], runs[1][1]), "run-to-run reproducible"


def test_forced_streaming_cg_small_rows(ctx):
    """Evolution matrix (23 entries per row on average): T = 8 lanes per row leave lanes idle, rows of
    every length from 5 to 100; Jacobi CG to convergence against the oracle."""
    _, ops = workload("bowl_mixing")
    A = (ops["M"] + 0.05 * (ops["Kh"] + ops["Kv"])).tocsr()
    b = np.random.default_rng(1).uniform(-1, 1, A.shape[0])
    dinv = 1.0 / A.diagonal()
    xo, so = krylov.cg(A, b, x0=np.zeros(b.size), M=dinv, atol=1e-10, rtol=1e-10)
    os.environ["NUPGCM_RESIDENT"] = "0"
    try:
        x = ctx.vector(b.size)
        st, hist = lib.cg_solve(ctx.csr(A), ctx.vector(b), x, dinv=ctx.vector(dinv), atol=1e-10, rtol=1e-10, history=512)
    finally:
        os.environ.pop("NUPGCM_RESIDENT")
    assert st.solved and abs(st.niter - so.niter) <= 1
    assert rel(x.download(), xo) < 1e-9


@pytest.mark.parametrize("orth,name", [(lib.ORTH_MGS, "mgs"), (lib.ORTH_CGS2_FUSED, "cgs2f")])
def test_streaming_vector_forms_with_many_rows_per_cta(ctx, orth, name):
    """5 300 rows per CTA (24 CTAs): more rows per thread than the register-resident Arnoldi forms hold, so
    the streaming forms run — with the first rows of the new vector parked in the idle footprint arena."""
    ops = refined_ops()
    A = ops["A"]
    y = ops["B"] @ ops["b_init"] + ops["b0"]
    M = np.full(y.size, ops["pscale"])
    xo, so = krylov.gmres(A, y, x0=np.zeros(y.size), M=M, atol=0.0, rtol=1e-30, memory=20, itmax=45, orth=name)
    os.environ["NUPGCM_GRID"] = "24"
    try:
        x = ctx.vector(y.size)
        st, hist = lib.gmres_solve(ctx.csr(A, drop_zeros=True), ctx.vector(y), x, pscale=ops["pscale"],
                                   atol=0.0, rtol=1e-30, itmax=45, memory=20, orth=orth, history=64)
    finally:
        os.environ.pop("NUPGCM_GRID")
    assert st.niter == 45
    assert np.allclose(hist, so.residuals, rtol=1e-9)
    assert rel(x.download(), xo) < 1e-9
