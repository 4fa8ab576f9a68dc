"""Set-up-time assembly of the operands of the solve path (host side, NumPy/SciPy).

Weak forms restated from reference ``src/inversion.jl:172-249`` and
``src/evolution.jl:209-296``; the assembly itself (done by Gridap in the reference) lives in
``gridap_lite``.  Row = test function, column = trial function; every (test, trial) DOF pair
sharing a cell is stored, numerically zero or not, exactly as Gridap's symbolic assembly does
(32 % of the stored 3-D inversion entries are explicit zeros, SURVEY.md finding 8).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .dofs import FEData
from .gridap_lite import restrict, restrict_vector
from .inputs import Forcings, Parameters, SurfaceDirichletBC, SurfaceFluxBC


def _scaled(c, s):
    if callable(c):
        return lambda x: s * np.asarray(c(x), dtype=np.float64)
    return s * float(c)


def build_A_inversion(fe_data: FEData, params: Parameters, ν):
    """LHS of the PG inversion (inversion.jl:133-148, forms :172-192), un-permuted, N x N CSR.

    Constant ν: ∫ α²ε² ν ∇u⊙∇v − (∇·v)p + q(∇·u) + f (ẑ×u)·v.
    Variable ν: the friction term is 2α²ε² ν σ(u)⊙σ(v) (inversion.jl:172-182).
    """
    dΩ = fe_data.mesh.dΩ
    U, P = fe_data.spaces.U, fe_data.spaces.P
    a2e2 = params.α ** 2 * params.ε ** 2
    lap = dΩ.matrix("stiff", U, U, coef=_scaled(ν, a2e2))
    cor = dΩ.matrix("mass", U, U, coef=params.f)
    uu = {(0, 0): lap, (1, 1): lap, (2, 2): lap, (0, 1): -cor, (1, 0): cor}
    if callable(ν):
        # 2 ν σ(u):σ(v) = ν (∇φᵢ·∇φⱼ δ_ab + ∂_b φᵢ ∂_a φⱼ)
        for a in range(3):
            for b in range(3):
                cross = dΩ.matrix("dd", U, U, coef=_scaled(ν, a2e2), comp=(b, a))
                uu[(a, b)] = uu[(a, b)] + cross if (a, b) in uu else cross
        uu = {k: _with_pattern(v, lap) for k, v in uu.items()}
    A_uu, _ = restrict(lap, U, U, uu)
    # −(∇·v) p : test u component a, trial p
    g = [dΩ.matrix("grad_test", U, P, coef=-1.0, comp=a) for a in range(3)]
    A_up, _ = restrict(g[0], U, P, {(a, 0): g[a] for a in range(3)})
    # q (∇·u) : test p, trial u component b
    d = [dΩ.matrix("grad_trial", P, U, coef=1.0, comp=b) for b in range(3)]
    A_pu, _ = restrict(d[0], P, U, {(0, b): d[b] for b in range(3)})
    A = sp.bmat([[A_uu, A_up], [A_pu, None]], format="coo")
    # bmat keeps the stored zeros of its blocks (COO concatenation); tocsr does not drop them
    A = A.tocsr()
    A.sort_indices()
    return A


def _with_pattern(m, pattern):
    """Re-express ``m`` on the (super-)pattern of ``pattern`` with explicit zeros."""
    z = pattern.copy()
    z.data[:] = 0.0
    coo = sp.coo_matrix((np.concatenate([z.data, m.tocoo().data]),
                         (np.concatenate([z.tocoo().row, m.tocoo().row]),
                          np.concatenate([z.tocoo().col, m.tocoo().col]))), shape=m.shape)
    out = coo.tocsr()
    out.sort_indices()
    return out


def build_B_inversion(fe_data: FEData, params: Parameters):
    """RHS matrix ``B`` (N x nb): ∫ α⁻¹ b (ẑ·v) (inversion.jl:199-219).  Also returns the
    Dirichlet part ``B_fd`` used for the lift in ``build_b_inversion``."""
    dΩ = fe_data.mesh.dΩ
    U, Bs = fe_data.spaces.U, fe_data.spaces.B
    m = dΩ.matrix("mass", U, Bs, coef=1.0 / params.α)
    B_ff, B_fd = restrict(m, U, Bs, {(2, 0): m})
    N = fe_data.spaces.nu + fe_data.spaces.np
    pad = sp.csr_matrix((N - B_ff.shape[0], B_ff.shape[1]))
    B = sp.vstack([B_ff, pad], format="csr")
    B.sort_indices()
    return B, B_fd


def build_b_inversion(fe_data: FEData, params: Parameters, forcings: Forcings, B_fd=None):
    """RHS vector b₀ (N): ∫_Γ α(τˣ v₁ + τʸ v₂) + ∫ α⁻¹ b_diri (ẑ·v) (inversion.jl:226-249)."""
    spaces = fe_data.spaces
    U = spaces.U
    N = spaces.nu + spaces.np
    out = np.zeros(N)
    own = np.zeros((U.n_owners, 3))
    dΓ = fe_data.mesh.dΓ
    for comp, τ in ((0, forcings.τˣ), (1, forcings.τʸ)):
        if callable(τ) or float(τ) != 0.0:
            if dΓ is None:
                raise ValueError("wind stress given but the mesh has no surface facets")
            fun = τ if callable(τ) else (lambda x, v=float(τ): np.full(x.shape[0], v))
            own[:, comp] = params.α * dΓ.vector(U, fun)
    out[:spaces.nu] = restrict_vector(own, U)
    if B_fd is None:
        _, B_fd = build_B_inversion(fe_data, params)
    out[:spaces.nu] += B_fd @ spaces.b_diri
    return out


def build_inversion_system(fe_data: FEData, params: Parameters, forcings: Forcings):
    """``A, B, b`` of the inversion (inversion.jl:121-126), un-permuted."""
    A = build_A_inversion(fe_data, params, forcings.ν)
    B, B_fd = build_B_inversion(fe_data, params)
    b = build_b_inversion(fe_data, params, forcings, B_fd)
    return A, B, b


# ------------------------------------------------------------------------------------------
# evolution
# ------------------------------------------------------------------------------------------

def build_matrix_vector(kind, fe_data: FEData, coef=1.0, dirs=(0, 1, 2)):
    """Matrix on free buoyancy DOFs and Dirichlet-lift vector ``a(b_diri, d)``
    (evolution.jl:256-260)."""
    Bs = fe_data.spaces.B
    m = fe_data.mesh.dΩ.matrix(kind, Bs, Bs, coef=coef, dirs=dirs)
    A_ff, A_fd = restrict(m, Bs, Bs, {(0, 0): m})
    return A_ff, A_fd @ fe_data.spaces.b_diri


def build_M(fe_data):
    return build_matrix_vector("mass", fe_data)                                 # evolution.jl:209-212


def build_Kh(fe_data, κₕ):
    return build_matrix_vector("stiff", fe_data, coef=κₕ, dirs=(0, 1))          # :224-227


def build_Kv(fe_data, κᵥ):
    return build_matrix_vector("stiff", fe_data, coef=κᵥ, dirs=(2,))            # :243-246


def build_rhs_diff(params: Parameters, fe_data: FEData, κᵥ):
    """∫ −N² κᵥ ∂z d (evolution.jl:269-278)."""
    Bs = fe_data.spaces.B
    v = fe_data.mesh.dΩ.vector("grad", Bs, coef=_scaled(κᵥ, -params.N2), comp=2)
    return restrict_vector(v.reshape(-1, 1), Bs)


def build_rhs_flux(params: Parameters, forcings: Forcings, fe_data: FEData):
    """∫_Γ α F d for a flux BC, zero for a Dirichlet BC (evolution.jl:280-296)."""
    bc = forcings.b_surface_bc
    Bs = fe_data.spaces.B
    if isinstance(bc, SurfaceDirichletBC):
        return np.zeros(Bs.nfree)
    if isinstance(bc, SurfaceFluxBC):
        flux = bc.flux if callable(bc.flux) else (lambda x, v=float(bc.flux): np.full(len(x), v))
        v = params.α * fe_data.mesh.dΓ.vector(Bs, flux)
        return restrict_vector(v.reshape(-1, 1), Bs)
    raise TypeError(type(bc))
