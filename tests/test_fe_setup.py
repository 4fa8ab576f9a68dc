"""Host-side set-up (gridap_lite + _forms) against the reference's fixtures.

Pins: the exact pin is the 2-D inversion matrix fixture of
reference test/bowl_mixing_tests.jl:59-62 (`A ≈ file["A_inversion"]`, rtol sqrt(eps)); DOF counts
and the discrete divergence of the golden velocity pin the 3-D numbering."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import golden, workload
from nupgcm_b200.gridap_lite import quadrature


def test_quadrature_exactness():
    import itertools
    import math

    def exact(a):
        d = len(a) - 1
        return math.factorial(d) * np.prod([math.factorial(k) for k in a]) / math.factorial(d + sum(a))

    for (bary, w), deg in [(quadrature.triangle(), 4), (quadrature.tetrahedron("keast11"), 4),
                           (quadrature.tetrahedron("keast15"), 5), (quadrature.segment(), 5)]:
        nb = bary.shape[1]
        for a in itertools.product(range(deg + 1), repeat=nb):
            if sum(a) > deg:
                continue
            assert abs((w * np.prod(bary ** np.array(a), axis=1)).sum() - exact(a)) < 1e-15


def test_inversion_matrix_matches_reference_fixture_2d():
    g = golden("A_bowl_mixing_2D.npz")
    m, n = g["A_inversion_shape"]
    ref = sp.csc_matrix((g["A_inversion_nzval"], g["A_inversion_rowval"], g["A_inversion_colptr"]),
                        shape=(m, n)).tocsr()
    ref.sort_indices()
    w, ops = workload("bowl_mixing", dim=2)
    d = w.fe_data().dofs
    A = ops["A"][d.inv_p_inversion][:, d.inv_p_inversion].tocsr()     # A[iperm, iperm], as the test does
    A.sort_indices()
    assert A.shape == ref.shape and A.nnz == ref.nnz == 38712           # stored entries incl. explicit zeros
    assert np.array_equal(A.indptr, ref.indptr) and np.array_equal(A.indices, ref.indices)
    assert abs(A - ref).max() <= 1e-15                                   # measured 5.6e-17
    # Julia's `≈` on matrices: norm(A − B) <= sqrt(eps) * max(norm(A), norm(B))
    assert np.linalg.norm(A.data - ref.data) <= np.sqrt(np.finfo(float).eps) * np.linalg.norm(ref.data)
    # and the reference's permutation check is only about consistency (bowl_mixing_tests.jl:60)
    assert sorted(g["iperm"].tolist()) == list(range(1098))


@pytest.mark.parametrize("dim,h,expect", [(2, 0.1, (990, 108, 349)), (3, 0.1, (14792, 1154, 5864)),
                                          (3, 0.08, (29353, 2042, 11211))])
def test_dof_counts(dim, h, expect):
    from nupgcm_b200 import workloads as W
    fe = W.bowl_mixing(dim=dim, h=h).fe_data()
    assert (fe.dofs.nu, fe.dofs.np, fe.dofs.nb) == expect


def test_surface_flux_has_no_buoyancy_dirichlet():
    from nupgcm_b200 import workloads as W
    assert W.bowl_surface_flux().fe_data().dofs.nb == 7434


def test_inversion_matrix_3d_sizes_and_golden_divergence():
    w, ops = workload("bowl_mixing")
    A = ops["A"]
    assert A.shape == (15946, 15946) and A.nnz == 1154824
    d = w.fe_data().dofs
    g = golden("bowl_mixing_3D.npz")
    x = np.concatenate([g["u"], np.zeros(d.np)])[d.p_inversion]
    div = (A @ x)[d.nu:]                                             # ∫ ψ_i ∇·u_golden
    assert np.abs(div).max() < 1e-16                                  # measured ~7e-19


def test_permutations_are_consistent():
    w, _ = workload("bowl_mixing")
    d = w.fe_data().dofs
    for p, ip in ((d.p_u, d.inv_p_u), (d.p_p, d.inv_p_p), (d.p_b, d.inv_p_b),
                  (d.p_inversion, d.inv_p_inversion)):
        assert np.array_equal(p[ip], np.arange(p.size))
    assert np.array_equal(d.p_inversion[:d.nu], d.p_u)
    assert np.array_equal(d.p_inversion[d.nu:], d.nu + d.p_p)


def test_evolution_matrices_share_pattern_and_are_symmetric():
    _, ops = workload("bowl_mixing")
    for k in ("M", "Kh", "Kv"):
        assert abs(ops[k] - ops[k].T).max() < 1e-15
        assert np.array_equal(ops[k].indices, ops["M"].indices)
    assert ops["M"].shape == (5864, 5864) and ops["M"].nnz == 132824
    # mass matrix integrates constants: sum(M_full) = volume; on free DOFs it is smaller
    assert 0 < ops["M"].sum() < 0.7854


def test_msh_and_npz_meshes_agree():
    import os
    src = "/root/reference/meshes/bowl3D_1.000000e-01_5.000000e-01.msh"
    if not os.path.exists(src):
        pytest.skip("reference checkout not present (GPU box)")
    from nupgcm_b200.gridap_lite import RawMesh, read_msh
    from nupgcm_b200.workloads import mesh_path
    a, b = read_msh(src), RawMesh.load_npz(mesh_path(3, 0.1))
    assert np.array_equal(a.nodes, b.nodes)
    for d in range(4):
        assert np.array_equal(a.elements[d], b.elements[d])
        assert a.element_names[d] == b.element_names[d]


def test_box_mesh_and_channel_basin_workload():
    """BASELINE config 4 on its declared substitute mesh: geometry, boundary names, DOF counts,
    and a 12-step oracle run of the production configuration (adaptive BDF1 + convection + eddy
    parameterisations) — Δt adapts, the eddy viscosity rebuild after step 10 changes the flow."""
    from nupgcm_b200 import workloads as W
    from nupgcm_b200.gridap_lite import box_mesh
    from oracle.stepping import cpu_model_for
    raw = box_mesh(3, 4, 2, z=(-0.125, 0.0))
    assert raw.nodes.shape == (4 * 5 * 3, 3) and raw.elements[3].shape == (3 * 4 * 2 * 6, 4)
    assert len(raw.elements[2]) == 2 * 2 * (3 * 4 + 3 * 2 + 4 * 2)          # two triangles per boundary quad
    assert sum(n == ("surface",) for n in raw.element_names[2]) == 2 * 3 * 4
    assert len(raw.elements[1]) == 2 * (3 + 4)                                # rim of the top face
    w = W.channel_basin_box()
    fe = w.fe_data()
    assert fe.mesh.dΩ.meas.sum() == pytest.approx(1.0 * 2.0 * 0.125, rel=1e-12)
    assert set(fe.mesh.model.tag_names) == {"bottom", "coastline", "surface"}
    d = fe.dofs
    assert d.np == 7 * 13 * 5 - 1 and d.nb == 13 * 25 * 9                    # flux BC: every P2 node is free
    ops = W.host_operands(w)
    cpu = cpu_model_for(w, ops, solver="direct")
    cpu.run(n_steps=12)
    nu = ops["nu"]
    assert np.isfinite(cpu.xu).all() and np.abs(cpu.xu[:nu]).max() < 1.0
    assert cpu.dts[0] == pytest.approx(0.05 * ops["tables"]["h_cells"].min() / 0.01)   # flow at rest
    assert cpu.dts[1] < 0.1 * cpu.dts[0]                                      # CFL takes over
    assert cpu.dts[11] > 2.0 * cpu.dts[10]                                    # ν_eddy rebuilt after step 10
