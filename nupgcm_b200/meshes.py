"""``Mesh``: discrete model plus the quadrature data of ``Measure(Ω, 4)`` / ``Measure(Γ, 4)``.

Mirrors reference ``src/meshes.jl:29-39`` (``Mesh``), ``:49-67`` (``get_p_t``), ``:94-108``
(``all_edges``) and ``:127-134`` (``compute_h_cells``).
"""
from __future__ import annotations

import numpy as np

from .gridap_lite import CellIntegrator, DiscreteModel, FacetIntegrator, RawMesh, read_msh


class Mesh:
    def __init__(self, ifile, degree: int = 4, surface_tags=("surface",), tet_rule=None):
        if degree != 4:
            raise ValueError("only the reference's default degree=4 measures are provided")
        if isinstance(ifile, RawMesh):
            raw = ifile
        elif str(ifile).endswith(".npz"):
            raw = RawMesh.load_npz(ifile)
        else:
            raw = read_msh(ifile)
        self.model = DiscreteModel(raw)
        self.dΩ = CellIntegrator(self.model, tet_rule=tet_rule)
        self.surface_tags = tuple(surface_tags)
        have = [t for t in self.surface_tags if t in self.model.tag_names]
        self.dΓ = FacetIntegrator(self.model, have) if have else None

    @property
    def dim(self):
        return self.model.dim


def get_p_t(model: DiscreteModel):
    """Node coordinates ``p`` and (0-based) connectivities ``t`` (meshes.jl:49-63)."""
    return model.nodes[:, :3], model.cells


def all_edges(t):
    """Unique edges of the *triangle-style* local edges (1,2),(2,3),(3,1) of each cell.

    The reference applies this triangle routine to tetrahedra as well (meshes.jl:94-96 called
    from inversion.jl:44-45), so only three of a tet's six edges enter; replicated on purpose.
    """
    e = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]], axis=0)
    e = np.sort(e, axis=1)
    return np.unique(e, axis=0)


def median_edge_length(model: DiscreteModel):
    """``h`` of the scalar inversion preconditioner (inversion.jl:44-48): lower median."""
    p, t = get_p_t(model)
    edges = all_edges(t)
    hs = np.sort(np.linalg.norm(p[edges[:, 0]] - p[edges[:, 1]], axis=1))
    return float(hs[len(hs) // 2 - 1])      # Julia hs[length(hs) ÷ 2], 1-based


def compute_h_cells(mesh: Mesh):
    """Longest edge of each cell (meshes.jl:127-134)."""
    x = mesh.model.nodes[mesh.model.cells]
    n = x.shape[1]
    h = np.zeros(x.shape[0])
    for i in range(n):
        for j in range(i + 1, n):
            h = np.maximum(h, np.linalg.norm(x[:, i] - x[:, j], axis=1))
    return h
