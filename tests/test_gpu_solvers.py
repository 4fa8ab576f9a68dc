"""Krylov parity: the persistent CUDA solvers against the CPU restatement of Krylov.jl
(iteration counts, residual histories) and against SciPy direct solves (relative residual
<= 1e-10, solution <= 1e-8 when run at tight tolerances).

Iteration-count criterion.  CG and short GMRES runs must agree with the oracle within ±1.  For
the 3-D inversion (thousands of GMRES(20) cycles with the scalar preconditioner) ±1 is not a
property any two implementations can have: the restarted recurrence amplifies rounding
differences — the oracle itself, fed a right-hand side perturbed by 1e-16 relative, moves from
5479 to 5161 iterations and its residual history departs by 1e-15 / 1e-7 / 1e-2 at iterations
100 / 1000 / 2000 (DESIGN.md "Parity", measured with tools/gmres_sensitivity.py).
There the test demands (i) residual histories equal to 1e-6 relative over the first 300
iterations, (ii) the count within the oracle's own perturbation spread (10 %), (iii) the
solution within the solver tolerance."""
GMRES_LONG_RUN_SPREAD = 0.10

import numpy as np
import pytest
import scipy.sparse.linalg as spla

from conftest import workload
from nupgcm_b200 import lib
from oracle import krylov

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _evol(ops, θ=0.05):
    return (ops["M"] + θ * (ops["Kh"] + ops["Kv"])).tocsr()


def test_cg_parity_default_tolerances(ctx):
    _, ops = workload("bowl_mixing")
    A = _evol(ops)
    rng = np.random.default_rng(0)
    b = rng.uniform(-1, 1, A.shape[0])
    dinv = 1.0 / A.diagonal()
    xo, so = krylov.cg(A, b, x0=np.zeros(b.size), M=dinv, atol=1e-6, rtol=1e-6)
    dA = ctx.csr(A)
    x = ctx.vector(b.size)
    st, hist = lib.cg_solve(dA, ctx.vector(b), x, dinv=ctx.vector(dinv), atol=1e-6, rtol=1e-6, history=4096)
    assert st.solved and so.solved
    assert abs(st.niter - so.niter) <= 1
    n = min(len(hist), len(so.residuals))
    assert np.allclose(hist[:n], so.residuals[:n], rtol=1e-6)
    assert rel(x.download(), xo) < 1e-6


def test_cg_tight_matches_direct(ctx):
    _, ops = workload("bowl_mixing")
    A = _evol(ops)
    rng = np.random.default_rng(1)
    b = rng.uniform(-1, 1, A.shape[0])
    x0 = rng.uniform(-1, 1, A.shape[0])            # non-trivial warm start
    x = ctx.vector(x0)
    st, _ = lib.cg_solve(ctx.csr(A), ctx.vector(b), x, dinv=ctx.vector(1.0 / A.diagonal()), atol=0.0, rtol=1e-14)
    got = x.download()
    assert st.solved
    assert np.linalg.norm(b - A @ got) / np.linalg.norm(b) < 1e-10
    assert rel(got, spla.spsolve(A.tocsc(), b)) < 1e-8
    # warm start from the converged answer: nothing to do
    st2, _ = lib.cg_solve(ctx.csr(A), ctx.vector(b), x, dinv=ctx.vector(1.0 / A.diagonal()), atol=1e-8, rtol=0.0)
    assert st2.niter == 0 and st2.solved
    # run-to-run reproducibility
    xa, xb = ctx.vector(x0), ctx.vector(x0)
    lib.cg_solve(ctx.csr(A), ctx.vector(b), xa, dinv=ctx.vector(1.0 / A.diagonal()))
    lib.cg_solve(ctx.csr(A), ctx.vector(b), xb, dinv=ctx.vector(1.0 / A.diagonal()))
    assert np.array_equal(xa.download(), xb.download())


def test_cg_itmax_and_zero_rhs(ctx):
    _, ops = workload("bowl_mixing")
    A = _evol(ops)
    b = np.random.default_rng(2).uniform(-1, 1, A.shape[0])
    x = ctx.vector(b.size)
    st, hist = lib.cg_solve(ctx.csr(A), ctx.vector(b), x, dinv=ctx.vector(1.0 / A.diagonal()), atol=0.0, rtol=1e-30, itmax=3, history=16)
    assert st.niter == 3 and not st.solved and len(hist) == 4      # non-convergence is not an error
    st, _ = lib.cg_solve(ctx.csr(A), ctx.vector(b.size), ctx.vector(b.size))
    assert st.niter == 0 and st.solved


@pytest.mark.parametrize("orth", [lib.ORTH_MGS, lib.ORTH_CGS2, lib.ORTH_CGS2_FUSED])
def test_gmres_parity_2d(ctx, orth):
    _, ops = workload("bowl_mixing", dim=2)
    A = ops["A"]
    rng = np.random.default_rng(3)
    b = rng.uniform(-1, 1, A.shape[0])
    M = np.full(b.size, ops["pscale"])
    xo, so = krylov.gmres(A, b, x0=np.zeros(b.size), M=M, atol=0.0, rtol=1e-6, memory=20, itmax=200000)
    x = ctx.vector(b.size)
    st, hist = lib.gmres_solve(ctx.csr(A), ctx.vector(b), x, pscale=ops["pscale"], atol=0.0, rtol=1e-6,
                               itmax=200000, memory=20, orth=orth, history=1 << 16)
    assert so.solved and st.solved
    tol = 1 if orth == lib.ORTH_MGS else max(2, so.niter // 200)
    assert abs(st.niter - so.niter) <= tol, (st.niter, so.niter)
    n = min(len(hist), len(so.residuals), 400)
    assert np.allclose(hist[:n], so.residuals[:n], rtol=1e-5)
    got = x.download()
    assert np.linalg.norm(b - A @ got) / np.linalg.norm(b) < 2e-6
    assert rel(got, xo) < 1e-4


def test_gmres_parity_3d_inversion(ctx):
    """Config-1-sized inversion (N = 15 946) with the reference's defaults (GMRES(20),
    P = I/h³, atol = rtol = 1e-6), cold start."""
    w, ops = workload("bowl_example", h=0.1)
    A = ops["A"]
    y = ops["B"] @ ops["b_init"] + ops["b0"]
    M = np.full(y.size, ops["pscale"])
    xo, so = krylov.gmres(A, y, x0=np.zeros(y.size), M=M, atol=1e-6, rtol=1e-6, memory=20)
    for drop in (False, True):
        x = ctx.vector(y.size)
        st, hist = lib.gmres_solve(ctx.csr(A, drop_zeros=drop), ctx.vector(y), x, pscale=ops["pscale"],
                                   atol=1e-6, rtol=1e-6, memory=20, history=1 << 16)
        assert st.solved == so.solved
        assert abs(st.niter - so.niter) <= GMRES_LONG_RUN_SPREAD * so.niter, (st.niter, so.niter)
        got = x.download()
        assert rel(got, xo) < 1e-4
        assert np.allclose(hist[:300], np.array(so.residuals)[:300], rtol=1e-6)
        # the stopping measure is ‖M r‖ with M = I/h³: check it directly on the answer
        assert ops["pscale"] * np.linalg.norm(y - A @ got) <= 1.01 * (1e-6 + 1e-6 * so.residuals[0])


def test_gmres_tight_matches_direct(ctx):
    w, ops = workload("bowl_mixing", dim=2)
    A = ops["A"]
    rng = np.random.default_rng(4)
    b = rng.uniform(-1, 1, A.shape[0])
    x = ctx.vector(b.size)
    st, _ = lib.gmres_solve(ctx.csr(A), ctx.vector(b), x, pscale=ops["pscale"], atol=0.0, rtol=1e-12,
                            itmax=2000000, memory=20)
    got = x.download()
    assert st.solved
    assert np.linalg.norm(b - A @ got) / np.linalg.norm(b) < 1e-10
    assert rel(got, spla.spsolve(A.tocsc(), b)) < 1e-8


def test_gmres_itmax_memory_and_warm_start(ctx):
    _, ops = workload("bowl_mixing", dim=2)
    A = ops["A"]
    b = np.random.default_rng(5).uniform(-1, 1, A.shape[0])
    M = np.full(b.size, ops["pscale"])
    for mem, itmax in ((20, 47), (5, 12), (1, 3)):
        xo, so = krylov.gmres(A, b, x0=np.zeros(b.size), M=M, atol=0.0, rtol=1e-30, memory=mem, itmax=itmax)
        x = ctx.vector(b.size)
        st, hist = lib.gmres_solve(ctx.csr(A), ctx.vector(b), x, pscale=ops["pscale"], atol=0.0,
                                   rtol=1e-30, itmax=itmax, memory=mem, history=256)
        assert st.niter == so.niter == itmax and not st.solved
        assert len(hist) == itmax + 1
        assert rel(x.download(), xo) < 1e-9
        assert np.allclose(hist, so.residuals, rtol=1e-9)
    # warm start: second call continues from the first call's answer
    x = ctx.vector(b.size)
    dA, db = ctx.csr(A), ctx.vector(b)
    lib.gmres_solve(dA, db, x, pscale=ops["pscale"], atol=0.0, rtol=1e-30, itmax=40)
    st, hist = lib.gmres_solve(dA, db, x, pscale=ops["pscale"], atol=0.0, rtol=1e-30, itmax=40, history=8)
    xo1, _ = krylov.gmres(A, b, x0=np.zeros(b.size), M=M, atol=0.0, rtol=1e-30, itmax=40)
    xo2, so2 = krylov.gmres(A, b, x0=xo1, M=M, atol=0.0, rtol=1e-30, itmax=40)
    assert rel(x.download(), xo2) < 1e-8
    assert np.isclose(hist[0], so2.residuals[0], rtol=1e-8)
    with pytest.raises(lib.NupgcmError, match="memory"):
        lib.gmres_solve(dA, db, x, memory=21)


def test_block_preconditioner_apply_and_gmres(ctx):
    """BlockDiagonalPreconditioner (preconditioners.jl:53-125) with CgPreconditioner blocks (:5-37):
    one application against the oracle, then GMRES(20) preconditioned with it (inversion.jl:60)
    against the oracle's restatement.  The inner CG on the friction block stops at its cap of 100
    iterations, so the operator is only approximately linear and outer counts are compared within
    5 %; the early residual history and the answer are compared tightly."""
    import nupgcm_b200 as npg
    from nupgcm_b200.preconditioners import block_operands
    w, ops = workload("bowl_mixing", dim=2)
    fe = w.fe_data()
    F, T = block_operands(w.params, fe)
    assert F.shape[0] == fe.dofs.nu and T.shape[0] == fe.dofs.np
    A = ops["A"].tocsr()
    rng = np.random.default_rng(2)
    b = rng.uniform(-1, 1, A.shape[0])
    arch = npg.GPU(0)
    bdp = npg.BlockDiagonalPreconditioner(arch, blocks=(F, T))
    Mo = krylov.BlockDiagonalPreconditioner(F, T)
    y = ctx.vector(b.size)
    bdp.mul_(y, ctx.vector(b))
    yo = Mo(b)
    assert rel(y.download(), yo) < 1e-7
    assert bdp.handle.info()["applies"] == 1 and bdp.handle.info()["inner_iters"] == Mo.inner_iters
    # fresh preconditioners (their warm starts are part of the operator)
    bdp = npg.BlockDiagonalPreconditioner(arch, blocks=(F, T))
    Mo = krylov.BlockDiagonalPreconditioner(F, T)
    xo, so = krylov.gmres(A, b, x0=np.zeros(b.size), M=Mo, atol=1e-6, rtol=1e-6, memory=20)
    x = ctx.vector(b.size)
    st, hist = lib.gmres_solve_prec(ctx.csr(A), bdp.handle, ctx.vector(b), x, atol=1e-6, rtol=1e-6,
                                    memory=20, history=4096)
    assert so.solved and st.solved
    assert abs(st.niter - so.niter) <= max(2, 0.05 * so.niter), (st.niter, so.niter)
    assert np.allclose(hist[:20], so.residuals[:20], rtol=1e-6)
    got = x.download()
    assert rel(got, xo) < 1e-3
    # far fewer outer iterations than with the scalar preconditioner (1829 on this system)
    assert st.niter < 1000
    # the toolkit path: InversionToolkit(arch, A, P=BlockDiagonalPreconditioner, B, b)
    inv = npg.InversionToolkit(arch, ops["A"], npg.BlockDiagonalPreconditioner(arch, blocks=(F, T)), ops["B"],
                               ops["b0"], drop_zeros=False)
    inv.solver.y.upload(b)
    npg.iterative_solve_(inv.solver)
    assert inv.solver.stats.solved and abs(inv.solver.stats.niter - so.niter) <= max(2, 0.05 * so.niter)
