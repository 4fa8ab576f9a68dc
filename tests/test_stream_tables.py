"""CPU tests of the streaming-SpMV table builder (csrc/csr.cu: build_stream_tables): the host walker
`nupgcm_diag_stream_spmv_host` traverses tiles / arena-staged footprints / per-warp streams of
jagged-diagonal slices exactly as the persistent kernels do and must reproduce SciPy's product,
writing every row exactly once and never reusing an arena slot before its dependency."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import workload
from nupgcm_b200 import lib


def _check(A, grid, arena=6656, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1, 1, A.shape[1])
    y, info = lib.stream_spmv_host(A, x, grid=grid, arena=arena)
    ref = A @ x
    scale = np.abs(A) @ np.abs(x) + 1e-300
    assert np.max(np.abs(y - ref) / scale) < 1e-14
    # padding: stream alignment (<= 7 per warp and CTA), <= 7 in front of every blocked slice, one tail piece
    assert info["entries"] >= A.nnz and info["entries"] < 1.005 * A.nnz + 8 * 11 * grid + 600
    return info


@pytest.mark.parametrize("grid,arena", [(148, 6656), (16, 1024), (3, 2048), (1, 1600)])
def test_inversion_matrix_2d(grid, arena):
    _, ops = workload("bowl_mixing", dim=2)
    A = ops["A"].tocsr().copy()
    A.eliminate_zeros()
    info = _check(A, grid, arena)
    assert info["tiles"] >= min(grid, A.shape[0] // 352)


def test_inversion_matrix_3d_many_tiles_and_bank_conflicts():
    _, ops = workload("bowl_mixing")
    A = ops["A"].tocsr().copy()
    A.eliminate_zeros()
    p = lib.rcm_order(A)
    P = A[p][:, p].tocsr()
    i1 = _check(P, 148)
    i2 = _check(P, 4, arena=4096)          # few CTAs: a dozen tiles per CTA go round the arena several times
    assert i2["tiles"] >= 4 * 11 and i1["tiles"] >= 148
    # bank-aware placement: the vector gathers of a position take close to the conflict-free 2 wavefronts
    assert i2["gather_wavefronts"] / i2["positions"] < 3.3      # random placement: about 6


def test_ragged_rows_long_rows_and_empty_rows():
    rng = np.random.default_rng(3)
    n = 1500
    rows = []
    for i in range(n):
        if i % 97 == 0:
            k = 0                                            # empty rows
        elif i % 50 == 1:
            k = int(rng.integers(97, 400))                   # a few rows far longer than the rest
        else:
            k = int(rng.integers(1, 60))
        lo = max(0, min(n - 500, i - 250))
        cols = np.sort(rng.choice(np.arange(lo, lo + 500), size=k, replace=False))
        rows.append(cols)
    indptr = np.concatenate([[0], np.cumsum([len(c) for c in rows])])
    indices = np.concatenate(rows).astype(np.int64)
    A = sp.csr_matrix((rng.uniform(-1, 1, len(indices)), indices, indptr), shape=(n, n))
    for grid, arena in [(5, 1024), (2, 1600), (40, 6656), (148, 2048)]:
        _check(A, grid, arena)


def test_tile_wider_than_the_arena_is_refused():
    n = 600
    A = sp.csr_matrix(np.ones((1, n))).tocsr()
    A = sp.vstack([A, sp.eye(n - 1, n, format="csr")]).tocsr()
    with pytest.raises(lib.NupgcmError):
        lib.stream_spmv_host(A, np.ones(n), grid=4, arena=256)
