"""Finite-element spaces for (u, p) and b.  Mirrors reference ``src/spaces.jl:31-72``."""
from __future__ import annotations

from .gridap_lite import LagrangeSpace
from .meshes import Mesh


class Spaces:
    """P2 vector velocity ``U``, P1 zero-mean pressure ``P`` and buoyancy ``B`` of order ``b_order``:
    2 (the reference's default, every test set-up and example) or 1 (the production runs,
    ``scratch/run.jl:152``: ``Spaces(mesh; ..., b_order=1)``).

    ``u_diri_vals`` must be zero vectors: the reference never lifts velocity Dirichlet data into
    the inversion right-hand side (``inversion.jl:226-249`` only lifts ``b``), so non-zero values
    would silently be ignored there; here they are rejected.
    """

    def __init__(self, mesh: Mesh, u_diri_tags=(), u_diri_masks=None, u_diri_vals=None,
                 b_diri_tags=(), b_diri_vals=None, u_order: int = 2, b_order: int = 2):
        if u_order != 2 or b_order not in (1, 2):
            raise ValueError("u_order must be 2 (Taylor-Hood P2-P1) and b_order 1 or 2")
        self.b_order = b_order
        model = mesh.model
        if u_diri_vals is not None:
            for v in u_diri_vals:
                if any(float(c) != 0.0 for c in v):
                    raise ValueError("non-zero velocity Dirichlet values are not supported")
        self.U = LagrangeSpace(model, 2, 3, u_diri_tags, u_diri_masks)
        self.P = LagrangeSpace(model, 1, 1, fix_last_owner=True)     # constraint=:zeromean
        self.B = LagrangeSpace(model, b_order, 1, b_diri_tags)
        if b_diri_vals is not None:
            self.B.set_dirichlet(b_diri_tags, b_diri_vals)
        # b_diri of the reference (spaces.jl:69): Dirichlet values with zero free values
        self.b_diri = self.B.dirichlet_values

    @property
    def nu(self):
        return self.U.nfree

    @property
    def np(self):
        return self.P.nfree

    @property
    def nb(self):
        return self.B.nfree
