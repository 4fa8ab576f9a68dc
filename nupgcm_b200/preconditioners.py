"""``CgPreconditioner`` / ``BlockDiagonalPreconditioner``.  Mirrors reference
``src/preconditioners.jl:5-37`` and ``:53-125``.

``BlockDiagonalPreconditioner(arch, params, fe_data)`` builds the two blocks the reference builds
(``:62-93``): ``P`` = the friction block of the inversion matrix with ν = 1 (``:75-81``; the
reference hard-codes ν = 1 with a warning) and ``T`` = pressure mass matrix / (α²ε²) with its
Jacobi diagonal (``:85-90``).  Each block's inverse is a ``CgPreconditioner``: CG to Krylov.jl's
default tolerance, warm-started from its previous answer, ``itmax`` = 100 for ``P`` (the
reference's GPU set-up, ``:106``) and 2n for ``T``.

Deviation (declared, see ``csrc/precond.cu``): the inner CG on ``P`` is Jacobi-preconditioned; the
reference's GPU set-up uses ILU(0) there.  The reference constructs this preconditioner nowhere
by default (``src/inversion.jl:60`` is commented out); pass it as ``P`` to ``InversionToolkit``.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import lib
from ._forms import build_A_inversion
from .architectures import GPU
from .dofs import FEData
from .inputs import Parameters


def block_operands(params: Parameters, fe_data: FEData):
    """Host operands of the two blocks in solver order: ``(P, T)`` as CSR matrices."""
    d = fe_data.dofs
    p = d.p_inversion
    A = build_A_inversion(fe_data, params, 1.0)                   # ν = 1 (preconditioners.jl:75-77)
    A0 = build_A_inversion(fe_data, params, 0.0)                  # everything but friction
    F = (A - A0)[p][:, p].tocsr()[:d.nu, :d.nu].tocsr()           # friction_only=true, [1:nu, 1:nu]
    F.eliminate_zeros()                                           # dropzeros!(A), :81
    F.sort_indices()
    M = fe_data.mesh.dΩ.matrix("mass", fe_data.spaces.P, fe_data.spaces.P)
    from .gridap_lite import restrict
    Mp, _ = restrict(M, fe_data.spaces.P, fe_data.spaces.P, {(0, 0): M})
    T = (Mp[d.p_p][:, d.p_p] / (params.α ** 2 * params.ε ** 2)).tocsr()
    T.sort_indices()
    return F, T


class BlockDiagonalPreconditioner:
    def __init__(self, arch, params: Parameters = None, fe_data: FEData = None, blocks=None,
                 P_itmax=100, T_itmax=0):
        if not isinstance(arch, GPU):
            raise NotImplementedError("nupgcm_b200 only provides the GPU() architecture")
        F, T = blocks if blocks is not None else block_operands(params, fe_data)
        ctx = arch.ctx
        self.P, self.T = ctx.csr(F), ctx.csr(T)
        self.P_dinv = ctx.vector(1.0 / F.diagonal())
        self.T_dinv = ctx.vector(1.0 / T.diagonal())
        self.handle = lib.BlockPrec(ctx, self.P, self.P_dinv, P_itmax, self.T, self.T_dinv, T_itmax)

    def mul_(self, y: lib.Vector, x: lib.Vector):
        """``mul!(y, bdp, x)`` (preconditioners.jl:118-125)."""
        return self.handle.apply(x, y)
