"""x-periodic channel of the channel_basin mesh (reference meshes/channel_basin_flat.jl:114-121: gmsh setPeriodic
of the east wall onto the west wall for y <= -L/4), reproduced on the structured substitute box by an
owner-level identification in gridap_lite (slave vertices / edges carry their masters' DOFs)."""
import numpy as np
import pytest

from nupgcm_b200 import workloads as W
from nupgcm_b200._forms import build_matrix_vector
from nupgcm_b200.gridap_lite import DiscreteModel, box_mesh
from nupgcm_b200.gridap_lite.fem import LagrangeSpace


def test_box_identification_counts_and_tags():
    nx, ny, nz = 4, 8, 3
    raw = box_mesh(nx, ny, nz, periodic_x_below=-0.5)
    m = DiscreteModel(raw)
    # nodes of the east face with y <= -0.5: (ny/4 + 1) rows of (nz + 1)
    nslave = (ny // 4 + 1) * (nz + 1)
    assert (m.vertex_master != np.arange(m.nv)).sum() == nslave
    x = m.nodes
    s = np.nonzero(m.vertex_master != np.arange(m.nv))[0]
    assert np.allclose(x[s][:, 0], 1.0) and np.allclose(x[m.vertex_master[s]][:, 0], 0.0)
    assert np.allclose(x[s][:, 1:], x[m.vertex_master[s]][:, 1:]) and x[s][:, 1].max() <= -0.5 + 1e-12
    # slave edges lie in the east face and map to the translated edge
    es = np.nonzero(m.edge_master != np.arange(m.ne))[0]
    mid = lambda e: 0.5 * (x[m.edges[e, 0]] + x[m.edges[e, 1]])                 # noqa: E731
    assert np.allclose(mid(es) - mid(m.edge_master[es]), [1.0, 0.0, 0.0])
    # the periodic faces are no boundary: no facet of the channel's side walls is left
    tri = x[m.boundary[2][0]]
    side = (np.isclose(tri[:, :, 0], 0.0).all(1) | np.isclose(tri[:, :, 0], 1.0).all(1)) & (tri[:, :, 1] <= -0.5 + 1e-12).all(1)
    assert not side.any()
    # DOF counts: P2 scalar space without Dirichlet conditions loses exactly the slave owners
    full = LagrangeSpace(DiscreteModel(box_mesh(nx, ny, nz)), 2, 1)
    per = LagrangeSpace(m, 2, 1)
    assert per.nfree == full.nfree - nslave - es.size
    assert np.array_equal(per.owner_dofs[s, 0], per.owner_dofs[m.vertex_master[s], 0])
    with pytest.raises(ValueError):
        box_mesh(2, 8, 3, periodic_x_below=-0.5)


@pytest.mark.parametrize("b_order", [2, 1])
def test_seam_is_interior(b_order):
    """On the periodic seam the weak Laplacian of g = y² + z sees no boundary: (K g)_i = ∫∇φ_i·∇g = −∫φ_i Δg
    = −2 (M 1)_i for every DOF whose support does not touch a wall, INCLUDING the master DOFs of the seam
    (with walls instead of the identification those rows carry the boundary term ∮ φ_i ∂ₙg)."""
    w = W.with_b_order(W.channel_basin_box(n=(4, 8, 3), periodic=True), b_order)
    fe = w.fe_data()
    Bs = fe.spaces.B
    K, _ = build_matrix_vector("stiff", fe)
    M, _ = build_matrix_vector("mass", fe)
    xo = Bs.owner_coordinates()
    g, _ = Bs.interpolate(lambda x: x[:, 1] ** 2 + x[:, 2])
    lhs = K @ g
    rhs = -2.0 * (M @ np.ones(Bs.nfree))
    # DOFs away from the true boundary: interior of the box, or on the seam strictly inside the channel
    own = np.nonzero(Bs.owner_is_master)[0]
    xm = xo[own]
    eps = 1e-9
    inside_yz = (xm[:, 1] > -1 + eps) & (xm[:, 2] > -0.125 + eps) & (xm[:, 2] < -eps)
    interior = inside_yz & (xm[:, 0] > eps) & (xm[:, 0] < 1 - eps) & (xm[:, 1] < 1 - eps)
    seam = inside_yz & (np.abs(xm[:, 0]) < eps) & (xm[:, 1] < -0.5 - eps)
    assert seam.sum() > 0 and interior.sum() > 0
    dof = Bs.owner_dofs[own, 0]
    if b_order == 2:            # P2 holds g exactly
        assert np.abs(lhs[dof[interior]] - rhs[dof[interior]]).max() < 1e-13
        assert np.abs(lhs[dof[seam]] - rhs[dof[seam]]).max() < 1e-13
    else:                       # P1 interpolant of y²: compare the seam rows with the row of the same (y, z) inside
        col = np.isclose(xm[:, 0], 0.5) & inside_yz & (xm[:, 1] < -0.5 - eps)
        key = lambda a: np.round(a[:, 1:] * 1e6).astype(np.int64)               # noqa: E731
        ks, kc = key(xm[seam]), key(xm[col])
        order_s, order_c = np.lexsort(ks.T), np.lexsort(kc.T)
        assert np.array_equal(ks[order_s], kc[order_c])
        assert np.abs(lhs[dof[seam]][order_s] - lhs[dof[col]][order_c]).max() < 1e-13
    # without the identification the same rows do carry a boundary term
    w2 = W.with_b_order(W.channel_basin_box(n=(4, 8, 3), periodic=False), b_order)
    fe2 = w2.fe_data()
    B2 = fe2.spaces.B
    K2, _ = build_matrix_vector("stiff", fe2)
    M2, _ = build_matrix_vector("mass", fe2)
    g2, _ = B2.interpolate(lambda x: x[:, 0] ** 2)          # ∂ₙg ≠ 0 on the east wall
    x2 = B2.owner_coordinates()
    wall = (np.abs(x2[:, 0] - 1.0) < eps) & (x2[:, 1] < -0.5 - eps) & (x2[:, 1] > -1 + eps) & (x2[:, 2] > -0.125 + eps) & (x2[:, 2] < -eps)
    d2 = B2.owner_dofs[wall, 0]
    assert np.abs((K2 @ g2)[d2] + 2.0 * (M2 @ np.ones(B2.nfree))[d2]).max() > 1e-3


def test_periodic_operands_are_translation_invariant():
    """With x-independent data the inversion right-hand side of the periodic channel has no x-dependence
    along the channel: rows of DOFs at the same (y, z) inside the channel agree, seam included."""
    w = W.channel_basin_box(n=(4, 8, 3), periodic=True)
    fe = w.fe_data()
    ops = W.host_operands(w)
    Bs = fe.spaces.B
    M = ops["M"]
    mass = np.empty(Bs.nfree)
    mass[fe.dofs.p_b] = M @ np.ones(Bs.nfree)                # back to Gridap order (M = M_gridap[p][:, p])
    own = np.nonzero(Bs.owner_is_master)[0]
    xo = Bs.owner_coordinates()[own]
    sel = (xo[:, 1] < -0.5 - 1e-9)                           # strictly inside the channel
    key = np.round(xo[sel][:, 1:] * 1e6).astype(np.int64)
    vals = mass[Bs.owner_dofs[own[sel], 0]]
    # vertex owners come in 4 x-translates (columns x = 0 — the seam — .25, .5, .75) with identical stars
    isv = own[sel] < fe.mesh.model.nv
    groups = {}
    for k, v in zip(map(tuple, key[isv]), vals[isv]):
        groups.setdefault(k, []).append(v)
    assert len(groups) >= 4
    for vs in groups.values():
        assert len(vs) == 4 and np.ptp(vs) < 1e-15            # identical lumped mass, seam column included
