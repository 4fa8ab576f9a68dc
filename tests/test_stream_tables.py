"""CPU tests of the streaming-SpMV table builder (csrc/csr.cu: build_stream_tables): the host walker
`nupgcm_diag_stream_spmv_host` traverses tiles / footprints / per-warp streams of jagged-diagonal
slices exactly as the persistent kernels do and must reproduce SciPy's product, writing every row exactly once."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import workload
from nupgcm_b200 import lib


def _check(A, grid, fmax, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1, 1, A.shape[1])
    y, nt, ne = lib.stream_spmv_host(A, x, grid=grid, fmax=fmax)
    ref = A @ x
    scale = np.abs(A) @ np.abs(x) + 1e-300
    assert np.max(np.abs(y - ref) / scale) < 1e-14
    assert ne >= A.nnz and ne < A.nnz + 8 * 11 * grid + 600     # only alignment padding + one tail piece
    return nt


@pytest.mark.parametrize("grid,fmax", [(148, 4096), (16, 512), (3, 256), (148, 300)])
def test_inversion_matrix_2d(grid, fmax):
    _, ops = workload("bowl_mixing", dim=2)
    A = ops["A"].tocsr().copy()
    A.eliminate_zeros()
    nt = _check(A, grid, fmax)
    assert nt >= min(grid, A.shape[0] // 64)


def test_inversion_matrix_3d_multi_tile():
    _, ops = workload("bowl_mixing")
    A = ops["A"].tocsr().copy()
    A.eliminate_zeros()
    p = lib.rcm_order(A)
    P = A[p][:, p].tocsr()
    nt1 = _check(P, 148, 4096)
    nt2 = _check(P, 8, 1024)          # few CTAs, small footprint cap: many tiles per CTA
    assert nt2 > 8 and nt1 >= 148


def test_ragged_rows_long_rows_and_empty_rows():
    rng = np.random.default_rng(3)
    n = 700
    rows = []
    for i in range(n):
        if i % 97 == 0:
            k = 0                                            # empty rows
        elif i % 50 == 1:
            k = int(rng.integers(97, 400))                   # a few rows far longer than the rest
        else:
            k = int(rng.integers(1, 60))
        cols = np.sort(rng.choice(n, size=min(k, n), replace=False))
        rows.append(cols)
    indptr = np.concatenate([[0], np.cumsum([len(c) for c in rows])])
    indices = np.concatenate(rows).astype(np.int64)
    A = sp.csr_matrix((rng.uniform(-1, 1, len(indices)), indices, indptr), shape=(n, n))
    for grid, fmax in [(5, 700), (2, 700), (40, 512), (148, 2048)]:
        _check(A, grid, fmax)


def test_row_wider_than_the_cap_is_refused():
    n = 600
    A = sp.csr_matrix(np.ones((1, n))).tocsr()
    A = sp.vstack([A, sp.eye(n - 1, n, format="csr")]).tocsr()
    with pytest.raises(lib.NupgcmError):
        lib.stream_spmv_host(A, np.ones(n), grid=4, fmax=256)
