"""``IterativeSolverToolkit`` and ``iterative_solve!``.  Mirrors reference
``src/iterative_solvers.jl:1-9,26-29`` (toolkit; ``x`` is the workspace's solution vector, so
every solve is warm-started from the previous answer) and ``:31-68`` (``iterative_solve!``).

The reference dispatches to a direct solve on the CPU (``:42-55``) and to
``Krylov.krylov_solve!`` on the GPU (``:58``); this package only has the latter, executed by one
persistent CUDA kernel per solve (``csrc/krylov.cu``)."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import lib


@dataclass
class SolverStats:
    """What the reference reads from ``workspace.stats`` (iterative_solvers.jl:61-63)."""
    solved: bool = False
    niter: int = 0
    timer: float = 0.0            # seconds on the device (CUDA events)
    residuals: np.ndarray = field(default_factory=lambda: np.zeros(0))
    inconsistent: bool = False
    rnorm: float = 0.0
    rnorm0: float = 0.0


class Diagonal:
    """``Diagonal(v)`` preconditioner; a scalar multiple of I is kept as a scalar
    (inversion.jl:54 builds ``(1/h^dim) I``)."""

    def __init__(self, v):
        if isinstance(v, lib.Vector):
            self.vec, self.scalar = v, None
        else:
            self.vec, self.scalar = None, float(v)


class IterativeSolverToolkit:
    def __init__(self, A, P: Diagonal, x: lib.Vector, y: lib.Vector, method: str, kwargs: dict,
                 label: str):
        self.A = A            # LHS matrix (reassigned at run time, evolution.jl:139-140)
        self.P = P            # preconditioner
        self.x = x            # solution vector == warm start of the next solve
        self.y = y            # RHS vector
        self.method = method  # "gmres" | "cg"  (the Krylov workspace type of the reference)
        self.kwargs = kwargs
        self.label = label
        self.stats = SolverStats()


def iterative_solve_(tk: IterativeSolverToolkit) -> IterativeSolverToolkit:
    """``iterative_solve!`` (iterative_solvers.jl:31-68), GPU branch."""
    kw = tk.kwargs
    hist = int(kw.get("history_cap", 65536)) if kw.get("history", True) else 0
    diag = tk.P if isinstance(tk.P, Diagonal) else Diagonal(1.0)
    common = dict(dinv=diag.vec, pscale=diag.scalar if diag.scalar is not None else 1.0,
                  atol=kw.get("atol", 1e-6), rtol=kw.get("rtol", 1e-6), itmax=kw.get("itmax", 0),
                  history=hist)
    if tk.method == "gmres" and hasattr(tk.P, "handle"):
        # operator preconditioner (BlockDiagonalPreconditioner, inversion.jl:60)
        st, res = lib.gmres_solve_prec(tk.A, tk.P.handle, tk.y, tk.x, atol=common["atol"], rtol=common["rtol"],
                                       itmax=common["itmax"], memory=kw.get("memory", 20), history=hist)
    elif tk.method == "gmres":
        if not kw.get("restart", True):
            raise NotImplementedError("restart=false GMRES is not provided")
        st, res = lib.gmres_solve(tk.A, tk.y, tk.x, memory=kw.get("memory", 20),
                                  orth=kw.get("orth", lib.ORTH_MGS), **common)
    elif tk.method == "cg":
        st, res = lib.cg_solve(tk.A, tk.y, tk.x, **common)
    else:
        raise ValueError(tk.method)
    tk.stats = SolverStats(solved=bool(st.solved), niter=int(st.niter), timer=st.device_ms * 1e-3,
                           residuals=res, inconsistent=bool(st.inconsistent), rnorm=st.rnorm,
                           rnorm0=st.rnorm0)
    return tk
