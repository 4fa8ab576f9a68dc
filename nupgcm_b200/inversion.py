"""``InversionToolkit`` and ``invert!``.  Mirrors reference ``src/inversion.jl:1-5`` (struct),
``:27-72`` (set-up: assemble, permute, scalar preconditioner, upload), ``:74-94`` (GMRES
workspace, ``memory=20``) and ``:101-110`` (``invert!``)."""
from __future__ import annotations

import numpy as np

from . import lib
from ._forms import build_inversion_system
from .architectures import GPU, on_architecture
from .dofs import FEData
from .inputs import Forcings, Parameters
from .iterative_solvers import Diagonal, IterativeSolverToolkit, iterative_solve_
from .meshes import median_edge_length


def permuted_inversion_system(fe_data: FEData, params: Parameters, forcings: Forcings):
    """Host operands in solver (RCM-permuted) order: ``A[p,p]``, ``B[p,:]`` (columns in the
    permuted buoyancy order), ``b[p]`` and the scalar preconditioner ``1/h^dim``
    (inversion.jl:34-54).

    The reference keeps ``B``'s columns in Gridap order because it multiplies the host
    ``b.free_values``; here buoyancy lives on the device in solver order, so the columns are
    permuted with ``p_b`` once at set-up instead of un-permuting ``b`` every step."""
    A, B, b = build_inversion_system(fe_data, params, forcings)
    p = fe_data.dofs.p_inversion
    A = A[p][:, p].tocsr()
    B = B[p][:, fe_data.dofs.p_b].tocsr()
    A.sort_indices()
    B.sort_indices()
    b = b[p]
    h = median_edge_length(fe_data.mesh.model)
    return A, B, b, 1.0 / h ** fe_data.mesh.dim


class InversionToolkit:
    def __init__(self, arch, *args, atol=1e-6, rtol=1e-6, itmax=0, memory=20, history=True,
                 verbose=False, restart=True, orth=lib.ORTH_MGS, drop_zeros=True):
        """``InversionToolkit(arch, fe_data, params, forcings; kwargs...)`` or
        ``InversionToolkit(arch, A, P, B, b; kwargs...)`` with host operands (inversion.jl:27,74).

        Extra keywords of this implementation: ``orth`` (Arnoldi orthogonalisation variant) and
        ``drop_zeros`` (do not store Gridap's explicit zeros — 32 % of the entries — on the
        device; default on, valid while the value pattern is fixed, i.e. constant ν)."""
        if not isinstance(arch, GPU):
            raise NotImplementedError("nupgcm_b200 only provides the GPU() architecture; the CPU "
                                      "path is the reference's own (no fallback)")
        if len(args) == 3:
            A, B, b, pscale = permuted_inversion_system(*args)
        elif len(args) == 4:
            A, pscale, B, b = args
        else:
            raise TypeError("InversionToolkit(arch, fe_data, params, forcings) or (arch, A, P, B, b)")
        self.arch = arch
        self.B = on_architecture(arch, B, drop_zeros=True)      # zeros of B never change
        self.b = on_architecture(arch, b)
        A_dev = on_architecture(arch, A, drop_zeros=drop_zeros)
        if getattr(arch, "comm", None) is not None:
            A_dev.shard(arch.comm)                              # GMRES becomes collective over the ranks
        N = A.shape[0]
        y = arch.ctx.vector(N)
        x = arch.ctx.vector(N)                                  # workspace.x .= 0 (inversion.jl:85)
        kwargs = dict(atol=atol, rtol=rtol, itmax=itmax, history=history, verbose=verbose,
                      restart=restart, memory=memory, orth=orth)
        # P: the scalar 1/h^dim of inversion.jl:54, or an operator preconditioner
        # (BlockDiagonalPreconditioner, the alternative of inversion.jl:60)
        P = pscale if hasattr(pscale, "handle") else Diagonal(pscale)
        self.solver = IterativeSolverToolkit(A_dev, P, x, y, "gmres", kwargs, "Inversion")


def invert_(inversion: InversionToolkit, b: lib.Vector):
    """``invert!(inversion, b)`` (inversion.jl:101-110): ``y = B b + b₀`` then solve.
    ``b`` is the device buoyancy in solver order."""
    s = inversion.solver
    s.y.copy_from(inversion.b)
    inversion.B.spmv(b, s.y, alpha=1.0, beta=1.0)
    iterative_solve_(s)
    return inversion
