"""CPU restatement of the per-step advection right-hand side.  TEST INFRASTRUCTURE ONLY.

Restates ``assemble_vector(d -> advection_lform(...), B_test)`` of reference
``src/model.jl:269-273`` with the linear forms of ``src/model.jl:292-295`` (BDF1) and
``:297-300`` (BDF2), followed by ``rhs_adv[perm]`` (``:274``), and the RHS combine of ``:278``.
Gridap (which executes it in the reference) is not vendored; the quadrature rule is a parameter
(SURVEY.md App. D item 10).  Written independently of the CUDA kernel *and* of
``nupgcm_b200.gridap_lite`` (own basis tabulation) so that it can check both.
"""
from __future__ import annotations

import numpy as np

_EDGES = {4: [(0, 1), (0, 2), (1, 2), (0, 3), (1, 3), (2, 3)], 3: [(0, 1), (0, 2), (1, 2)]}


def _p2(bary):
    nq, nb = bary.shape
    edges = _EDGES[nb]
    val = np.empty((nq, nb + len(edges)))
    der = np.zeros((nq, nb + len(edges), nb))
    for i in range(nb):
        val[:, i] = bary[:, i] * (2.0 * bary[:, i] - 1.0)
        der[:, i, i] = 4.0 * bary[:, i] - 1.0
    for e, (i, j) in enumerate(edges):
        val[:, nb + e] = 4.0 * bary[:, i] * bary[:, j]
        der[:, nb + e, i] = 4.0 * bary[:, j]
        der[:, nb + e, j] = 4.0 * bary[:, i]
    return val, der


def _p1(bary):
    nq, nv = bary.shape
    return bary.copy(), np.broadcast_to(np.eye(nv), (nq, nv, nv)).copy()


def _b_basis(tables):
    """Buoyancy basis: P2 (reference default) or P1 (``Spaces(...; b_order=1)``, scratch/run.jl:152),
    told apart by the number of local DOFs in ``cell_b``."""
    bary = tables["bary"]
    return _p1(bary) if tables["cell_b"].shape[1] == bary.shape[1] else _p2(bary)


def rhs_adv(tables, scheme, dt, N2, b, b_prev, u, u_prev):
    """Advection RHS in solver (permuted) row order.

    ``tables``: dict with ``cell_b`` (nc, nloc) and ``cell_u`` (nc, nloc, 3) indices into the
    *extended* vectors (free DOFs in permuted order followed by Dirichlet values), ``grad``
    (nc, d+1, 3), ``vol`` (nc,), ``bary`` (nq, d+1), ``w`` (nq,), ``nb`` free count,
    ``b_dirichlet``, ``u_dirichlet`` value arrays.  ``b, b_prev`` (nb) and ``u, u_prev`` (nu) are
    free values in permuted order.
    """
    cb, cu = tables["cell_b"], tables["cell_u"]
    phi, _ = _p2(tables["bary"])                                          # velocity: always P2
    phib, dphib = _b_basis(tables)
    grad = tables["grad"]
    wq = tables["w"][None, :] * tables["vol"][:, None]                    # (nc, nq)
    bx = np.concatenate([b, tables["b_dirichlet"]])
    bpx = np.concatenate([b_prev, tables["b_dirichlet"]])
    ux = np.concatenate([u, tables["u_dirichlet"]])
    upx = np.concatenate([u_prev, tables["u_dirichlet"]])
    if scheme == 2:          # model.jl:297-300
        bs = 2.0 * bx[cb] - bpx[cb]
        us = 2.0 * ux[cu] - upx[cu]
        lin = (4.0 / 3.0) * bx[cb] - (1.0 / 3.0) * bpx[cb]
        fac = (2.0 / 3.0) * dt
    elif scheme == 1:        # model.jl:292-295
        bs, us, lin, fac = bx[cb], ux[cu], bx[cb], dt
    else:
        raise ValueError("scheme must be 1 (BDF1) or 2 (BDF2)")
    gphi = np.einsum("qik,ckd->cqid", dphib, grad)                        # ∇φᵢ at q
    gb = np.einsum("cqid,ci->cqd", gphi, bs)                              # ∇b*
    uq = np.einsum("qi,cid->cqd", phi, us)                                # u*
    lq = np.einsum("qi,ci->cq", phib, lin)
    val = lq - fac * (np.einsum("cqd,cqd->cq", uq, gb) + uq[:, :, 2] * N2)
    fe = np.einsum("cq,cq,qi->ci", wq, val, phib)
    out = np.zeros(tables["nb"] + tables["b_dirichlet"].size)
    np.add.at(out, cb.ravel(), fe.ravel())
    return out[:tables["nb"]]


def cfl_dt(tables, u, cfl_factor=0.8, u_min=0.01):
    """``update_Δt!`` of reference ``src/timesteppers.jl:108-119``: c · min_K h_K / max(|u|_{L∞(K)},
    u_min) with |u|_{L∞(K)} the largest Euclidean speed over the cell's quadrature points and
    h_K = ``tables["h_cells"]`` (``compute_h_cells``, src/meshes.jl:127-134)."""
    phi, _ = _p2(tables["bary"])
    ux = np.concatenate([u, tables["u_dirichlet"]])
    uq = np.einsum("qi,cid->cqd", phi, ux[tables["cell_u"]])
    speed = np.sqrt((uq ** 2).sum(axis=2)).max(axis=1)
    return cfl_factor * float(np.min(tables["h_cells"] / np.maximum(speed, u_min)))


def kv_rebuild(tables, kv_q, α, N2, κc, N2min, b, pattern):
    """Convection parameterisation, reference ``src/model.jl:229-246``: κᵥ(x_q) = κᵥ⁰(x_q) +
    κᶜ(1 + tanh(−α(N² + ∂z b)/N²min))/2 (``src/inputs.jl:87-91``), then ``build_Kᵥ``
    (``src/evolution.jl:243-246``: Kᵥ = ∫κᵥ ∂z b ∂z d and the Dirichlet lift rhsᵥ = a(b_diri, d),
    ``:256-260``) and ``build_rhs_diff`` (``:269-278``: ∫ −N² κᵥ ∂z d), all in solver order.

    ``pattern``: SciPy CSR with the evolution sparsity pattern (values ignored).  Returns
    ``(Kv_csr, rhs_v, rhs_diff)``."""
    import scipy.sparse as sp
    cb = tables["cell_b"]
    nb = tables["nb"]
    _, dphi = _b_basis(tables)
    gz = tables["grad"][:, :, 2]                                          # (nc, d+1)
    dz = np.einsum("qik,ck->cqi", dphi, gz)                               # ∂z φ_i at q
    bx = np.concatenate([b, tables["b_dirichlet"]])
    bdx = np.concatenate([np.zeros(nb), tables["b_dirichlet"]])
    dzb = np.einsum("cqi,ci->cq", dz, bx[cb])
    dzd = np.einsum("cqi,ci->cq", dz, bdx[cb])
    kap = kv_q + κc * (1.0 + np.tanh(-(α * (N2 + dzb)) / N2min)) / 2.0
    wq = tables["w"][None, :] * tables["vol"][:, None] * kap
    ke = np.einsum("cq,cqi,cqj->cij", wq, dz, dz)
    fv = np.einsum("cq,cq,cqi->ci", wq, dzd, dz)
    fd = np.einsum("cq,cqi->ci", wq, dz) * (-N2)
    nloc = cb.shape[1]
    rows = np.repeat(cb, nloc, axis=1).ravel()
    cols = np.tile(cb, (1, nloc)).ravel()
    keep = (rows < nb) & (cols < nb)
    K = sp.coo_matrix((ke.ravel()[keep], (rows[keep], cols[keep])), shape=(nb, nb)).tocsr()
    # same stored pattern as the other evolution matrices
    P = sp.csr_matrix((np.zeros(pattern.nnz), pattern.indices, pattern.indptr), shape=pattern.shape)
    K = (K + P).tocsr()
    K.sort_indices()
    rv = np.zeros(nb + tables["b_dirichlet"].size)
    rd = np.zeros_like(rv)
    np.add.at(rv, cb.ravel(), fv.ravel())
    np.add.at(rd, cb.ravel(), fd.ravel())
    return K, rv[:nb], rd[:nb]


def nu_friction(tables, f_q, a2e2, α, N2, N2min, b, N, smoothing=10.0, ν_min=1.0):
    """Eddy parameterisation, reference ``src/model.jl:160-170``: ν(x_q) = ν_eddy(α(N² + ∂z b))
    (``src/inputs.jl:130-137``: f²/sqrt(N²min² + (α∂z b)²), LogSumExp-limited below by ν_min) and the
    friction block ∫ 2α²ε² ν σ(u)⊙σ(v) (``src/inversion.jl:172-182``) as an N x N CSR in solver
    order (rows/columns beyond the velocity DOFs are empty)."""
    import scipy.sparse as sp
    cb, cu = tables["cell_b"], tables["cell_u"]
    nb, nu = tables["nb"], tables["nu"]
    _, dphi = _p2(tables["bary"])
    g = np.einsum("qik,ckd->cqid", dphi, tables["grad"])                  # ∇φ_i at q (velocity, P2)
    bx = np.concatenate([b, tables["b_dirichlet"]])
    dzb = np.einsum("qik,ck,ci->cq", _b_basis(tables)[1], tables["grad"][:, :, 2], bx[cb])
    αbz = α * (N2 + dzb)
    ν = f_q * (f_q / np.sqrt(N2min ** 2 + αbz * αbz))
    ν = np.logaddexp(smoothing * ν_min, smoothing * ν) / smoothing
    wq = a2e2 * tables["w"][None, :] * tables["vol"][:, None] * ν
    S = np.einsum("cq,cqid,cqjd->cij", wq, g, g)
    T = np.einsum("cq,cqja,cqib->ciajb", wq, g, g)                        # ∂_a φ_j ∂_b φ_i
    blk = T + np.einsum("cij,ab->ciajb", S, np.eye(3))
    rows = np.broadcast_to(cu[:, :, :, None, None], blk.shape).ravel()
    cols = np.broadcast_to(cu[:, None, None, :, :], blk.shape).ravel()
    keep = (rows < nu) & (cols < nu)
    return sp.coo_matrix((blk.ravel()[keep], (rows[keep], cols[keep])), shape=(N, N)).tocsr()


def rhs_combine(rhs_adv_v, θ, dt, rhs_diff, rhs_flux, rhs_m, rhs_h, rhs_v):
    """``y = rhs_adv + θ rhs_diff + Δt rhs_flux − (rhsₘ + θ (rhsₕ + rhsᵥ))`` (model.jl:278)."""
    return rhs_adv_v + θ * rhs_diff + dt * rhs_flux - (rhs_m + θ * (rhs_h + rhs_v))
