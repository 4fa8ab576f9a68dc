"""A C-language caller of the ABI (tests/capi/test_abi.c, compiled with gcc against
include/nupgcm_b200.h): 1-based int64 CSR in, one call per solve, results checked against values the
CPU oracle wrote.  The CPU part only compiles and links it (the header must be plain C); the GPU part
runs it."""
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import ROOT, workload
from oracle import krylov

SRC = os.path.join(ROOT, "tests", "capi", "test_abi.c")
LIBDIR = os.path.join(ROOT, "nupgcm_b200")


def _build(tmp_path):
    exe = str(tmp_path / "test_abi")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-O1", "-I", os.path.join(ROOT, "include"), SRC,
           "-o", exe, "-L", LIBDIR, "-lnupgcm_b200", "-lm", f"-Wl,-rpath,{LIBDIR}"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return exe


def test_c_caller_compiles_and_links(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True)          # no argument: usage, before any device call
    assert out.returncode == 2 and "usage" in out.stderr


@pytest.mark.gpu
def test_c_caller_matches_oracle(tmp_path):
    _, ops = workload("bowl_mixing", dim=2)
    # one SPD system for both solvers (so that CG is legal): the 2-D evolution matrix
    A = sp.csr_matrix(ops["M"] + 0.05 * (ops["Kh"] + ops["Kv"]))
    A.sort_indices()
    n = A.shape[0]
    y = np.random.default_rng(0).uniform(-1, 1, n)
    dinv = 1.0 / A.diagonal()
    tol, pscale, itmax = 1e-10, 3.0, 45
    x_cg, s_cg = krylov.cg(A, y, x0=np.zeros(n), M=dinv, atol=tol, rtol=tol)
    x_gm, s_gm = krylov.gmres(A, y, x0=np.zeros(n), M=np.full(n, pscale), atol=0.0, rtol=1e-30, memory=20, itmax=itmax)
    path = tmp_path / "system.bin"
    with open(path, "wb") as f:
        np.array([n, A.nnz, itmax, s_cg.niter], dtype=np.int64).tofile(f)
        np.array([pscale, tol], dtype=np.float64).tofile(f)
        (A.indptr.astype(np.int64) + 1).tofile(f)
        (A.indices.astype(np.int64) + 1).tofile(f)
        A.data.astype(np.float64).tofile(f)
        for v in (y, dinv, x_cg, np.asarray(s_gm.residuals[:itmax + 1], dtype=np.float64), x_gm):
            np.asarray(v, dtype=np.float64).tofile(f)
    assert len(s_gm.residuals) >= itmax + 1
    out = subprocess.run([_build(tmp_path), str(path)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "test_abi ok" in out.stdout
