"""CPU restatement of the Krylov.jl recurrences nuPGCM invokes.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference legs may
import this module; the product (``nupgcm_b200``) must never route through it.

PARITY UNPINNED: the algorithm lives in Krylov.jl 0.10.6 (``Manifest.toml:686-690``), which is not
vendored under ``/root/reference`` and cannot be executed here (no Julia).  The reference holds no
fixture recording iteration counts or residual histories.  The restatement follows
Krylov.jl's published CG and GMRES as called from
``src/iterative_solvers.jl:58`` with the keyword sets of ``src/inversion.jl:76,89`` (GMRES,
``memory=20, restart=true, atol=rtol=1e-6, itmax=0 -> 2n``) and ``src/evolution.jl:60-64,125``
(CG), and the left preconditioners of ``src/inversion.jl:54`` (``(1/h^dim) I``) and
``src/evolution.jl:149,167`` (``1/diag(A)``).  It is validated against SciPy direct solves.

Conventions: ``A`` is a SciPy CSR matrix, ``M`` a vector (diagonal left preconditioner, applied
by multiplication, ``ldiv=false``) or ``None``; ``x0`` is the warm start (the reference aliases
``x`` to ``workspace.x`` so every solve is warm-started from the previous answer,
``iterative_solvers.jl:26-29``).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

EPS = np.finfo(np.float64).eps


@dataclass
class Stats:
    niter: int = 0
    solved: bool = False
    inconsistent: bool = False
    residuals: list = field(default_factory=list)
    status: str = "unknown"


def _apply(M, r):
    if M is None:
        return r.copy()
    return M(r) if callable(M) else M * r


class BlockDiagonalPreconditioner:
    """Reference ``src/preconditioners.jl:53-125`` with ``CgPreconditioner`` blocks (``:5-37``):
    ``y[:n1] = CG(P)⁻¹ x[:n1]``, ``y[n1:] = CG(T)⁻¹ x[n1:]``; each inner CG uses Krylov.jl's default
    tolerances (atol = rtol = sqrt(eps)), an iteration cap and the block's previous answer as its
    initial guess (``cgp.workspace.x``, ``:25``).  Inner preconditioners are Jacobi (the GPU set-up
    of the reference uses ILU(0) for ``P``, ``:102-107`` — same deviation as the device code)."""

    def __init__(self, P, T, P_itmax=100, T_itmax=0):
        self.P, self.T = P.tocsr(), T.tocsr()
        self.Pd, self.Td = 1.0 / self.P.diagonal(), 1.0 / self.T.diagonal()
        self.n1 = self.P.shape[0]
        self.P_itmax, self.T_itmax = P_itmax, T_itmax
        self.xp, self.xt = np.zeros(self.n1), np.zeros(self.T.shape[0])
        self.inner_iters = 0

    def __call__(self, x):
        self.xp, s1 = cg(self.P, x[:self.n1], x0=self.xp, M=self.Pd, itmax=self.P_itmax, history=False)
        self.xt, s2 = cg(self.T, x[self.n1:], x0=self.xt, M=self.Td, itmax=self.T_itmax, history=False)
        self.inner_iters += s1.niter + s2.niter
        return np.concatenate([self.xp, self.xt])


def cg(A, b, x0=None, M=None, atol=np.sqrt(EPS), rtol=np.sqrt(EPS), itmax=0, history=True):
    """Krylov.jl ``cg!`` with a left (Jacobi) preconditioner and warm start.

    Stopping measure is sqrt(rᵀ M r) against ``atol + rtol * (initial measure)``.
    """
    n = A.shape[0]
    st = Stats()
    dx = np.zeros(n) if x0 is None else np.array(x0, dtype=np.float64)
    x = np.zeros(n)
    r = b - A @ dx if x0 is not None else np.array(b, dtype=np.float64)
    z = _apply(M, r)
    p = z.copy()
    γ = float(r @ z)
    rnorm = np.sqrt(γ)
    if history:
        st.residuals.append(rnorm)
    if γ == 0.0:
        st.solved = True
        st.status = "x = 0 is a zero-residual solution"
        return x + dx, st
    it = 0
    if itmax == 0:
        itmax = 2 * n
    pnorm2 = γ
    ε = atol + rtol * rnorm
    solved = rnorm <= ε
    tired = it >= itmax
    zero_curv = False
    while not (solved or tired or zero_curv):
        Ap = A @ p
        pAp = float(p @ Ap)
        if pAp <= EPS * pnorm2 and abs(pAp) <= EPS * pnorm2:
            zero_curv = True
            st.inconsistent = True
            continue
        α = γ / pAp
        x += α * p
        r -= α * Ap
        z = _apply(M, r)
        γn = float(r @ z)
        rnorm = np.sqrt(γn)
        if history:
            st.residuals.append(rnorm)
        solved = (rnorm <= ε) or (rnorm + 1.0 <= 1.0)
        if not solved:
            β = γn / γ
            pnorm2 = γn + β * β * pnorm2
            γ = γn
            p = z + β * p
        it += 1
        tired = it >= itmax
    st.niter = it
    st.solved = solved
    st.status = ("solution good enough given atol and rtol" if solved else
                 "zero curvature detected" if zero_curv else "maximum number of iterations exceeded")
    return x + dx, st


def sym_givens(a, b):
    """Krylov.jl ``sym_givens`` for reals: (c, s, ρ) with [c s; s −c][a; b] = [ρ; 0]."""
    if b == 0.0:
        c = 1.0 if a == 0.0 else np.sign(a)
        return c, 0.0, abs(a)
    if a == 0.0:
        return 0.0, np.sign(b), abs(b)
    if abs(b) > abs(a):
        t = a / b
        s = np.sign(b) / np.sqrt(1.0 + t * t)
        c = s * t
        return c, s, b / s
    t = b / a
    c = np.sign(a) / np.sqrt(1.0 + t * t)
    s = c * t
    return c, s, a / c


def gmres(A, b, x0=None, M=None, atol=np.sqrt(EPS), rtol=np.sqrt(EPS), itmax=0, memory=20,
          restart=True, history=True, orth="mgs"):
    """Krylov.jl ``gmres!``: restarted, left-preconditioned, modified Gram-Schmidt.

    ``orth``: ``"mgs"`` is what Krylov.jl does (sequential dot/axpy pairs); ``"cgs2"`` (classical
    Gram-Schmidt applied twice) is provided to measure how far a batched-reduction variant moves
    the iteration count; ``"cgs2f"`` is CGS2 with two instead of three global synchronisations
    per iteration (what ``NUPGCM_ORTH_CGS2_FUSED`` runs on the device): the norm of the new vector
    comes from the second projection by Pythagoras, ‖q₂‖² = ‖q₁‖² − ‖h₂‖², and the next operator
    application is done on q₁ (the vector that was exchanged); by the Arnoldi relation that adds
    V_{k+1}(H̄_k h₂)/H — a vector inside the span the next projection removes — so only the next
    Hessenberg column needs the correction.  Identical to CGS2 in exact arithmetic.
    """
    if not restart:
        raise NotImplementedError("the reference always passes restart=true (inversion.jl:76)")
    n = A.shape[0]
    st = Stats()
    mem = memory
    x = np.zeros(n)
    if x0 is not None:
        dx = np.array(x0, dtype=np.float64)
        w = b - A @ dx
        x += dx
    else:
        w = np.array(b, dtype=np.float64)
    r0 = _apply(M, w)
    β = float(np.linalg.norm(r0))
    rnorm = β
    if history:
        st.residuals.append(β)
    ε = atol + rtol * rnorm
    if β == 0.0:
        st.solved = True
        st.status = "x = 0 is a zero-residual solution"
        return x, st
    it = 0
    if itmax == 0:
        itmax = 2 * n
    inner_itmax = itmax
    btol = EPS ** 0.75
    breakdown = False
    solved = rnorm <= ε
    tired = it >= itmax
    npass = 0
    V = np.zeros((mem + 1, n))
    while not (solved or tired or breakdown):
        V[:] = 0.0
        c = np.zeros(mem)
        s = np.zeros(mem)
        R = np.zeros(mem * (mem + 1) // 2)
        z = np.zeros(mem + 1)
        nr = 0
        xr = np.zeros(n)
        if npass >= 1:
            w = b - A @ x
            r0 = _apply(M, w)
        β = float(np.linalg.norm(r0))
        z[0] = β
        V[0] = r0 / β
        npass += 1
        k = 0                          # inner_iter
        inner_tired = False
        raw, inv_h, corr = r0, 1.0 / β, None          # cgs2f: exchanged vector, its scale, H̄ h₂
        Hbar = np.zeros((mem + 1, mem))
        while not (solved or inner_tired or breakdown):
            k += 1
            if orth == "cgs2f":
                q = _apply(M, (A @ raw) * inv_h)       # = Â v_k + V_k c,  c = H̄ h₂ / H (Arnoldi relation)
            else:
                w = A @ V[k - 1]
                q = _apply(M, w)
            if orth == "cgs2f":
                h1 = V[:k] @ q
                q -= h1 @ V[:k]
                raw = q.copy()                        # q₁: what the other owners gather next
                h2 = V[:k] @ q
                ssq = float(q @ q)
                q -= h2 @ V[:k]
                # V_k c lies in the span just projected out: only the coefficients need the correction
                hk = h1 + h2 - (inv_h * corr[:k] if corr is not None else 0.0)
                R[nr:nr + k] = hk
                Hbar[:k, k - 1] = hk
                hsq = ssq - float(h2 @ h2)
            elif orth == "mgs":
                for i in range(k):
                    h = float(V[i] @ q)
                    R[nr + i] = h
                    q -= h * V[i]
            elif orth == "cgs2":
                h1 = V[:k] @ q
                q -= h1 @ V[:k]
                h2 = V[:k] @ q
                q -= h2 @ V[:k]
                R[nr:nr + k] = h1 + h2
            elif orth == "cgs":
                h1 = V[:k] @ q
                q -= h1 @ V[:k]
                R[nr:nr + k] = h1
            else:
                raise ValueError(orth)
            Hbis = float(np.sqrt(max(hsq, 0.0))) if orth == "cgs2f" else float(np.linalg.norm(q))
            if orth == "cgs2f":
                Hbar[k, k - 1] = Hbis
                corr = Hbar[:k + 1, :k] @ h2
                inv_h = 1.0 / Hbis if Hbis > 0 else 0.0
            for i in range(k - 1):
                tmp = c[i] * R[nr + i] + s[i] * R[nr + i + 1]
                R[nr + i + 1] = s[i] * R[nr + i] - c[i] * R[nr + i + 1]
                R[nr + i] = tmp
            c[k - 1], s[k - 1], R[nr + k - 1] = sym_givens(R[nr + k - 1], Hbis)
            ζ = s[k - 1] * z[k - 1]
            z[k - 1] = c[k - 1] * z[k - 1]
            rnorm = abs(ζ)
            if history:
                st.residuals.append(rnorm)
            nr += k
            solved = (rnorm <= ε) or (rnorm + 1.0 <= 1.0)
            breakdown = Hbis <= btol
            inner_tired = k >= min(mem, inner_itmax)
            if not (solved or inner_tired or breakdown):
                V[k] = q / Hbis
                z[k] = ζ
        # back substitution R y = z (packed upper triangular, column-major by column)
        y = z[:k].copy()
        for i in range(k, 0, -1):
            pos = nr + i - k - 1                    # 0-based position of r_{i,k}
            for j in range(k, i, -1):
                y[i - 1] -= R[pos] * y[j - 1]
                pos = pos - j + 1
            if abs(R[pos]) <= btol:
                y[i - 1] = 0.0
                st.inconsistent = True
            else:
                y[i - 1] /= R[pos]
        xr = y @ V[:k]
        x += xr
        inner_itmax -= k
        it += k
        tired = it >= itmax
    st.niter = it
    st.solved = solved
    st.status = ("solution good enough given atol and rtol" if solved else
                 "found approximate least-squares solution" if breakdown else
                 "maximum number of iterations exceeded")
    return x, st
