"""CPU study for SURVEY row a8 / f-3 (reference src/preconditioners.jl:53-125): how many OUTER GMRES(20) iterations the
block-diagonal preconditioner needs on a 3-D inversion system, depending on what preconditions the inner CG on the
friction block — Jacobi (what csrc/precond.cu does today), ILU(0) with exact triangular solves (the reference's GPU
set-up, `kp_ilu0`, :102-107), or ILU(0) whose triangular solves are replaced by s Jacobi sweeps (= a truncated Neumann
series: s SpMVs with the strict triangles, no dependent chains — the form a persistent kernel could run) — and what
that costs in SpMV-equivalents.  Runs the oracle (test infrastructure) on the host; nothing here is product code.

    python tools/block_prec_study.py [h]"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nupgcm_b200 import workloads as W                        # noqa: E402
from nupgcm_b200.preconditioners import block_operands        # noqa: E402
from oracle import krylov                                     # noqa: E402


def ilu0(A):
    """ILU(0) of a CSR matrix with sorted indices: returns (L strict lower, D = diag(U), U strict upper), A ≈ (I+L)(D+U)."""
    A = A.tocsr().copy()
    A.sort_indices()
    n = A.shape[0]
    ip, ix, a = A.indptr, A.indices, A.data
    diag = np.empty(n, dtype=np.int64)
    for i in range(n):
        diag[i] = ip[i] + np.searchsorted(ix[ip[i]:ip[i + 1]], i)
    for i in range(n):
        lo, hi = ip[i], ip[i + 1]
        cols_i = ix[lo:hi]
        for kk in range(lo, diag[i]):
            k = ix[kk]
            a[kk] /= a[diag[k]]
            # row_i[j] -= a_ik * row_k[j] for j > k in both patterns
            ks, ke = diag[k] + 1, ip[k + 1]
            if ks == ke:
                continue
            pos = np.searchsorted(cols_i, ix[ks:ke])
            ok = (pos < cols_i.size)
            ok[ok] &= cols_i[pos[ok]] == ix[ks:ke][ok]
            a[lo + pos[ok]] -= a[kk] * a[ks:ke][ok]
    M = sp.csr_matrix((a, ix, ip), shape=A.shape)
    return sp.tril(M, -1).tocsr(), M.diagonal(), sp.triu(M, 1).tocsr()


class Counter:
    def __init__(self):
        self.spmv_F = 0.0       # SpMV-equivalents with the friction block


class IluPrec:
    """z = (D+U)⁻¹ (I+L)⁻¹ r, exact (sweeps = 0) or with each triangular solve replaced by `sweeps` Jacobi sweeps."""

    def __init__(self, F, sweeps, cnt):
        self.L, self.d, self.U = ilu0(F)
        self.sweeps, self.cnt = sweeps, cnt
        self.Lfull = (sp.identity(F.shape[0], format="csr") + self.L).tocsr()
        self.Ufull = (sp.diags(self.d) + self.U).tocsr()
        self.Us = sp.diags(1.0 / self.d) @ self.U            # (D+U) = D (I + D⁻¹U)

    def __call__(self, r):
        if self.sweeps == 0:
            y = spla.spsolve_triangular(self.Lfull, r, lower=True, unit_diagonal=True)
            z = spla.spsolve_triangular(self.Ufull, y, lower=False)
            self.cnt.spmv_F += 1.0                            # same entries as one SpMV, but in dependent levels
            return z
        y = r.copy()
        for _ in range(self.sweeps):
            y = r - self.L @ y
        w = y / self.d
        z = w.copy()
        for _ in range(self.sweeps):
            z = w - self.Us @ z
        self.cnt.spmv_F += self.sweeps                        # 2 s products with half of the pattern each
        return z


class BlockPrec:
    def __init__(self, F, T, inner, cnt, P_itmax=100):
        self.F, self.T, self.inner, self.cnt = F, T, inner, cnt
        self.Td = 1.0 / T.diagonal()
        self.n1 = F.shape[0]
        self.xp, self.xt = np.zeros(self.n1), np.zeros(T.shape[0])
        self.P_itmax = P_itmax
        self.inner_iters = 0

    def __call__(self, x):
        self.xp, s1 = krylov.cg(self.F, x[:self.n1], x0=self.xp, M=self.inner, itmax=self.P_itmax, history=False)
        self.xt, s2 = krylov.cg(self.T, x[self.n1:], x0=self.xt, M=self.Td, history=False)
        self.inner_iters += s1.niter
        self.cnt.spmv_F += s1.niter + 1
        return np.concatenate([self.xp, self.xt])


def exact_blocks(F, T):
    luF, luT = spla.splu(F.tocsc()), spla.splu(T.tocsc())
    n1 = F.shape[0]
    return lambda x: np.concatenate([luF.solve(x[:n1]), luT.solve(x[n1:])])


def schwarz(h, parts_list=(148, 8), overlaps=(1, 2)):
    """Restricted additive Schwarz on the persistent solvers' own row blocks (the library's RCM order and nnz-balanced
    partition): a block = the CTA's rows + `overlap` layers of graph neighbours, inverted exactly."""
    from nupgcm_b200 import lib
    w = W.bowl_example(h=h)
    ops = W.host_operands(w)
    A = ops["A"].tocsr()
    A.eliminate_zeros()
    n = A.shape[0]
    y = ops["B"] @ ops["b_init"] + ops["b0"]
    p = lib.rcm_order(A)
    Ap, yp = A[p][:, p].tocsr(), y[p]
    res = lambda x: np.linalg.norm(Ap @ x - yp) / np.linalg.norm(yp)                  # noqa: E731
    x, s = krylov.gmres(Ap, yp, x0=np.zeros(n), M=np.full(n, ops["pscale"]), atol=1e-6, rtol=1e-6, memory=20)
    print(f"bowl3D h = {h:g} (N = {n}), scalar I/h^3: {s.niter} / {res(x):.1e}", flush=True)
    rp = Ap.indptr
    G = (abs(Ap) + abs(Ap).T).tocsr()
    for parts in parts_list:
        cost, total = rp[:-1] + 4.0 * np.arange(n), rp[n] + 4.0 * n                   # build_partition of csrc/csr.cu
        part = [0] + [int(np.searchsorted(cost, total * q / parts)) for q in range(1, parts)] + [n]
        for ov in overlaps:
            sets = []
            for i in range(parts):
                idx = np.arange(part[i], part[i + 1])
                for _ in range(ov):
                    idx = np.unique(np.concatenate([idx, G[idx].indices]))
                sets.append(idx)
            lus = [spla.splu(Ap[idx][:, idx].tocsc()) for idx in sets]
            own = [(idx >= part[i]) & (idx < part[i + 1]) for i, idx in enumerate(sets)]

            def ras(r):
                z = np.zeros(n)
                for idx, lu, o in zip(sets, lus, own):
                    z[idx[o]] = lu.solve(r[idx])[o]
                return z
            x, s = krylov.gmres(Ap, yp, x0=np.zeros(n), M=ras, atol=1e-6, rtol=1e-6, memory=20, itmax=3000)
            print(f"   {parts:3d} blocks, overlap {ov}: {s.niter:5d} / {res(x):.1e}   mean block {np.mean([len(i) for i in sets]):.0f} rows "
                  f"({n / parts:.0f} owned)", flush=True)


def main():
    """`--schwarz [h]`: restricted additive Schwarz on the solvers' row blocks.  Default: scalar preconditioner against the block preconditioner with EXACT inner solves (its best case) on the
    3-D bowl and on the channel_basin box.  `--inner [h]`: the inner-preconditioner variants on the bowl (slow)."""
    if "--schwarz" in sys.argv:
        rest = [a for a in sys.argv[1:] if a != "--schwarz"]
        return schwarz(float(rest[0]) if rest else 0.1)
    if "--inner" not in sys.argv:
        for name, w in (("bowl3D h = 0.1, examples/bowl_mixing.jl parameters (eps = 0.2, alpha = 0.5)", W.bowl_example(h=0.1)),
                        ("bowl2D h = 0.1 (bowl_mixing_tests.jl)", W.bowl_mixing(dim=2)),
                        ("channel_basin box 6x12x4, periodic channel, P1 buoyancy (eps = 0.32, alpha = 0.125)",
                         W.with_b_order(W.channel_basin_box(periodic=True), 1))):
            ops = W.host_operands(w)
            A = ops["A"].tocsr()
            n = A.shape[0]
            y = ops["B"] @ ops["b_init"] + ops["b0"]
            if not np.any(y):
                y = np.random.default_rng(2).uniform(-1, 1, n)
            F, T = block_operands(w.params, w.fe_data())
            x0, s0 = krylov.gmres(A, y, x0=np.zeros(n), M=np.full(n, ops["pscale"]), atol=1e-6, rtol=1e-6, memory=20)
            x1, s1 = krylov.gmres(A, y, x0=np.zeros(n), M=exact_blocks(F, T), atol=1e-6, rtol=1e-6, memory=20, itmax=20000)
            print(f"{name}: N = {n}\n    scalar P = I/h^3: {s0.niter} GMRES(20) iterations (solved={s0.solved});  block-diagonal with exact "
                  f"inner solves: {s1.niter} outer iterations (solved={s1.solved})", flush=True)
        return
    args = [a for a in sys.argv[1:] if a != "--inner"]
    h = float(args[0]) if args else 0.1
    w = W.bowl_example(h=h)
    ops = W.host_operands(w)
    A = ops["A"].tocsr()
    y = ops["B"] @ ops["b_init"] + ops["b0"]
    F, T = block_operands(w.params, w.fe_data())
    n = A.shape[0]
    print(f"bowl3D h = {h:g}: N = {n}, nnz(A) = {A.nnz}, friction block {F.shape[0]} rows / {F.nnz} non-zeros; cold start, rtol = atol = 1e-6")
    t0 = time.time()
    x, st = krylov.gmres(A, y, x0=np.zeros(n), M=np.full(n, ops["pscale"]), atol=1e-6, rtol=1e-6, memory=20)
    print(f"scalar P = I/h^3                      : {st.niter:7d} outer iterations = {st.niter:7d} SpMV(A)                       solved={st.solved}  ({time.time() - t0:.0f} s)", flush=True)
    fa = F.nnz / A.nnz
    for name, make in (("block, inner CG + Jacobi", lambda c: 1.0 / F.diagonal()),
                       ("block, inner CG + ILU(0) exact solves", lambda c: IluPrec(F, 0, c)),
                       ("block, inner CG + ILU(0), 4 sweeps", lambda c: IluPrec(F, 4, c))):
        cnt = Counter()
        t0 = time.time()
        inner = make(cnt)
        M = BlockPrec(F, T, inner, cnt)
        x2, s2 = krylov.gmres(A, y, x0=np.zeros(n), M=M, atol=1e-6, rtol=1e-6, memory=20, itmax=300)
        err = np.linalg.norm(x2 - x) / np.linalg.norm(x)
        total = s2.niter + cnt.spmv_F * fa
        print(f"{name:38s}: {s2.niter:7d} outer iterations (cap 300), {M.inner_iters:7d} inner CG iterations on the friction block "
              f"({M.inner_iters / max(s2.niter, 1):.1f} per apply) ≈ {total:9.0f} SpMV(A)-equivalents  solved={s2.solved}  "
              f"rnorm/rnorm0 {s2.residuals[-1] / s2.residuals[0]:.2e}  ({time.time() - t0:.0f} s)", flush=True)


if __name__ == "__main__":
    main()
