"""Element tables handed to ``nupgcm_mesh_create``: what the device needs to evaluate the
advection linear form of reference ``src/model.jl:292-300`` cell by cell.

Indices address *extended* vectors: free DOFs in solver (RCM-permuted) order followed by the
Dirichlet values, so that the kernel gathers straight from the solver vectors and no
``x[inv_perm]`` / ``rhs[perm]`` round trip (model.jl:274,282,312) is needed per step."""
from __future__ import annotations

import numpy as np

from .dofs import FEData
from .meshes import compute_h_cells


def _extended_index(owner_dofs, inv_perm, nfree):
    """Gridap DOF ids (>=0 free, <0 Dirichlet) -> positions in [free (permuted); Dirichlet]."""
    ids = owner_dofs
    out = np.empty(ids.shape, dtype=np.int64)
    free = ids >= 0
    out[free] = inv_perm[ids[free]]
    out[~free] = nfree + (-ids[~free] - 1)
    return out


def element_tables(fe_data: FEData) -> dict:
    sp = fe_data.spaces
    dofs = fe_data.dofs
    integ = fe_data.mesh.dΩ
    Bs, U = sp.B, sp.U
    b_ext = _extended_index(Bs.owner_dofs[:, 0], dofs.inv_p_b, Bs.nfree)
    u_ext = _extended_index(U.owner_dofs, dofs.inv_p_u, U.nfree)
    cell_b = b_ext[Bs.cell_owners]                       # (nc, nloc)
    cell_u = u_ext[U.cell_owners]                        # (nc, nloc, 3)
    return {
        "cell_b": cell_b.astype(np.int32), "cell_u": cell_u.astype(np.int32),
        "grad": np.ascontiguousarray(integ.grad), "vol": np.ascontiguousarray(integ.meas),
        "bary": np.ascontiguousarray(integ.bary), "w": np.ascontiguousarray(integ.w),
        "nb": Bs.nfree, "nu": U.nfree,
        "b_dirichlet": Bs.dirichlet_values.copy(), "u_dirichlet": U.dirichlet_values.copy(),
        "rule": integ.rule,
        "h_cells": compute_h_cells(fe_data.mesh),           # for update_Δt! (timesteppers.jl:108-119)
    }
