"""Parity of the CUDA kernels (through the C ABI) against the CPU oracle / SciPy.

Tolerances: SpMV and vector kernels <= 1e-13 relative (FP64, different summation order);
element RHS <= 1e-12 relative; everything must be bit-reproducible run to run."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import workload
from nupgcm_b200 import lib
from oracle.element_rhs import rhs_adv, rhs_combine

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def test_device_is_b200(ctx):
    info = ctx.device_info()
    assert info["cc"][0] == 10 and info["sm_count"] >= 100
    free, total = ctx.mem_status()
    assert 0 < free <= total


@pytest.mark.parametrize("which", ["A", "M", "B"])
@pytest.mark.parametrize("drop", [False, True])
def test_spmv_matches_scipy(ctx, which, drop):
    _, ops = workload("bowl_mixing")
    A = ops[which]
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, A.shape[1])
    y0 = rng.uniform(-1, 1, A.shape[0])
    dA = ctx.csr(A, drop_zeros=drop)
    info = dA.info()
    assert info["nnz_given"] == A.nnz
    assert info["nnz_stored"] == (np.count_nonzero(A.data) if drop else A.nnz)
    dx, dy = ctx.vector(x), ctx.vector(y0)
    dA.spmv(dx, dy)
    assert rel(dy.download(), A @ x) < 1e-13
    dy.upload(y0)
    dA.spmv(dx, dy, alpha=-0.5, beta=2.0)
    assert rel(dy.download(), -0.5 * (A @ x) + 2.0 * y0) < 1e-13
    # bit-reproducible
    a = dA.spmv(dx, ctx.vector(A.shape[0])).download()
    b = dA.spmv(dx, ctx.vector(A.shape[0])).download()
    assert np.array_equal(a, b)


def test_spmv_edge_cases(ctx):
    # empty rows, a single dense row, 1x1, and an all-zero matrix with drop_zeros
    rng = np.random.default_rng(3)
    m = sp.random(257, 129, density=0.02, random_state=4, format="lil")
    m[5, :] = rng.uniform(-1, 1, 129)
    m[7, :] = 0
    m = m.tocsr()
    x = rng.uniform(-1, 1, 129)
    y = ctx.csr(m).spmv(ctx.vector(x), ctx.vector(257)).download()
    assert rel(y, m @ x) < 1e-13
    one = sp.csr_matrix(np.array([[2.5]]))
    assert ctx.csr(one).spmv(ctx.vector([2.0]), ctx.vector(1)).download()[0] == 5.0
    z = sp.csr_matrix((np.zeros(3), np.array([0, 1, 2]), np.array([0, 1, 2, 3])), shape=(3, 3))
    dz = ctx.csr(z, drop_zeros=True)
    assert dz.info()["nnz_stored"] == 0
    assert np.array_equal(dz.spmv(ctx.vector([1.0, 2, 3]), ctx.vector(3)).download(), np.zeros(3))


def test_csr_rejects_bad_input(ctx):
    m = sp.identity(4, format="csr")
    bad = sp.csr_matrix((m.data, np.array([0, 1, 2, 9]), m.indptr), shape=(4, 4))
    bad.has_canonical_format = True
    with pytest.raises(lib.NupgcmError, match="column index"):
        lib.CsrMatrix(ctx, bad)
    with pytest.raises(lib.NupgcmError, match="length mismatch"):
        ctx.csr(m).spmv(ctx.vector(3), ctx.vector(4))


def test_vector_ops(ctx):
    rng = np.random.default_rng(1)
    for n in (1, 31, 1000, 100003):
        x, y = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
        dx, dy = ctx.vector(x), ctx.vector(y)
        assert abs(dx.dot(dy) - x @ y) <= 1e-13 * np.sqrt(n) * max(1.0, abs(x @ y))
        assert abs(dx.norm2() - np.linalg.norm(x)) <= 1e-13 * np.linalg.norm(x)
        dy.axpby(0.3, dx, -1.5)
        assert rel(dy.download(), 0.3 * x - 1.5 * y) < 1e-15
        m, nan = dx.maxabs()
        assert m == np.abs(x).max() and not nan
        m2, _ = dx.maxabs(max(1, n // 2))
        assert m2 == np.abs(x[:max(1, n // 2)]).max()
        z = ctx.vector(n)
        lib.diag_apply(z, dx, dy)
        assert np.array_equal(z.download(), x * dy.download())
    x[0] = np.nan
    assert ctx.vector(x).maxabs()[1]
    assert ctx.vector(0).norm2() == 0.0


def test_gather_and_fill(ctx):
    rng = np.random.default_rng(2)
    n = 5000
    x = rng.uniform(-1, 1, n)
    perm = rng.permutation(n)
    out = ctx.vector(n).gather_from(ctx.vector(x), ctx.index(perm)).download()
    assert np.array_equal(out, x[perm])
    out1 = ctx.vector(n).gather_from(ctx.vector(x), ctx.index(perm + 1, index_base=1)).download()
    assert np.array_equal(out1, x[perm])
    assert np.array_equal(ctx.vector(7).fill(2.5).download(), np.full(7, 2.5))
    with pytest.raises(lib.NupgcmError, match="exceeds source"):
        ctx.vector(3).gather_from(ctx.vector(2), ctx.index([0, 1, 2]))


def test_combine_and_inv_diag(ctx):
    _, ops = workload("bowl_mixing")
    M, Kh, Kv = (ctx.csr(ops[k]) for k in ("M", "Kh", "Kv"))
    A = ctx.csr(ops["M"])
    θ = 0.0123
    A.combine(M, Kh, Kv, θ)
    ref = (ops["M"] + θ * (ops["Kh"] + ops["Kv"])).tocsr()
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, ref.shape[0])
    assert rel(A.spmv(ctx.vector(x), ctx.vector(x.size)).download(), ref @ x) < 1e-13
    dinv = A.inv_diag(ctx.vector(x.size)).download()
    assert rel(dinv, 1.0 / ref.diagonal()) < 1e-15
    # update_values: with and without dropped zeros
    for drop in (False, True):
        B = ctx.csr(ops["A"], drop_zeros=drop)
        B.update_values(2.0 * ops["A"].data)
        xx = rng.uniform(-1, 1, ops["A"].shape[0])
        assert rel(B.spmv(ctx.vector(xx), ctx.vector(xx.size)).download(), 2.0 * (ops["A"] @ xx)) < 1e-13


@pytest.mark.parametrize("b_order", [2, 1])
@pytest.mark.parametrize("name", ["bowl_mixing", "bowl_dirichlet", "bowl_surface_flux"])
@pytest.mark.parametrize("scheme", [1, 2])
def test_element_rhs_matches_oracle(ctx, name, scheme, b_order):
    """b_order = 1: first-order buoyancy with P2 velocity (Spaces(...; b_order=1), scratch/run.jl:152)."""
    kw = {"dim": 2} if name == "bowl_mixing" and scheme == 1 else {}
    _, ops = workload(name, b_order=b_order, **kw)
    assert ops["tables"]["cell_b"].shape[1] == {2: ops["tables"]["cell_u"].shape[1], 1: ops["tables"]["bary"].shape[1]}[b_order]
    tb = ops["tables"]
    rng = np.random.default_rng(7)
    nb, nu = ops["nb"], ops["nu"]
    b, bp = rng.uniform(-1, 1, nb), rng.uniform(-1, 1, nb)
    u, up = rng.uniform(-1, 1, nu), rng.uniform(-1, 1, nu)
    N = nu + ops["np"]
    xu, xup = np.concatenate([u, rng.uniform(-1, 1, N - nu)]), np.concatenate([up, np.zeros(N - nu)])
    mesh = lib.ElementMesh(ctx, tb)
    out = ctx.vector(nb)
    mesh.rhs_adv(scheme, 0.1, 2.0, ctx.vector(b), ctx.vector(bp), ctx.vector(xu), ctx.vector(xup), out)
    ref = rhs_adv(tb, scheme, 0.1, 2.0, b, bp, u, up)
    got = out.download()
    assert rel(got, ref) < 1e-12
    mesh.rhs_adv(scheme, 0.1, 2.0, ctx.vector(b), ctx.vector(bp), ctx.vector(xu), ctx.vector(xup), out)
    assert np.array_equal(out.download(), got)               # deterministic gather order
    vs = [rng.uniform(-1, 1, nb) for _ in range(5)]
    y = ctx.vector(nb)
    lib.rhs_combine(y, out, 0.37, 0.1, *[ctx.vector(v) for v in vs])
    assert rel(y.download(), rhs_combine(got, 0.37, 0.1, *vs)) < 1e-15
