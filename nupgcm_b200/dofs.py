"""DOF permutations.  Mirrors reference ``src/dofs.jl:1-13,31-41`` (``DoFHandler``),
``:70-100`` (``compute_dof_perms``) and ``:104-124`` (``FEData``).

The reference calls ``CuthillMcKee.symrcm(M, true, false)`` (CuthillMcKee.jl 0.1.0, not vendored);
SciPy's reverse Cuthill-McKee gives an ordering of equal bandwidth but different tie-breaking
(SURVEY.md §8c), which is irrelevant downstream because every solver sees already-permuted
operands.  All permutations here are 0-based.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.sparse.csgraph import reverse_cuthill_mckee

from .meshes import Mesh
from .spaces import Spaces


def invperm(p):
    inv = np.empty_like(p)
    inv[p] = np.arange(p.size, dtype=p.dtype)
    return inv


def compute_dof_perm(M):
    """RCM of a (mass-matrix) sparsity graph (dofs.jl:98-100)."""
    pat = sp.csr_matrix((np.ones(M.nnz, dtype=np.int8), M.indices, M.indptr), shape=M.shape)
    return np.asarray(reverse_cuthill_mckee(pat, symmetric_mode=True), dtype=np.int64)


def _free_graph(space, cells_owners_pattern):
    """Sparsity graph over free DOFs of a space: all components of owners sharing a cell."""
    pat = cells_owners_pattern.tocoo()
    nc = space.ncomp
    r = space.owner_dofs[pat.row][:, :, None].repeat(nc, axis=2).ravel()
    c = space.owner_dofs[pat.col][:, None, :].repeat(nc, axis=1).ravel()
    ok = (r >= 0) & (c >= 0)
    g = sp.coo_matrix((np.ones(ok.sum(), dtype=np.int8), (r[ok], c[ok])),
                      shape=(space.nfree, space.nfree)).tocsr()
    return g


def owner_pattern(space):
    """Owner-level adjacency (owners sharing a cell)."""
    co = space.cell_owners
    n = co.shape[1]
    r = np.repeat(co, n, axis=1).ravel()
    c = np.tile(co, (1, n)).ravel()
    g = sp.coo_matrix((np.ones(r.size, dtype=np.int8), (r, c)),
                      shape=(space.n_owners, space.n_owners)).tocsr()
    g.data[:] = 1
    return g


def compute_dof_perms(spaces: Spaces):
    out = []
    for s in (spaces.U, spaces.P, spaces.B):
        out.append(compute_dof_perm(_free_graph(s, owner_pattern(s))))
    return tuple(out)


class DoFHandler:
    def __init__(self, p_u, p_p, p_b):
        self.p_u, self.p_p, self.p_b = (np.asarray(p, dtype=np.int64) for p in (p_u, p_p, p_b))
        self.nu, self.np, self.nb = self.p_u.size, self.p_p.size, self.p_b.size
        self.inv_p_u = invperm(self.p_u)
        self.inv_p_p = invperm(self.p_p)
        self.inv_p_b = invperm(self.p_b)
        self.p_inversion = np.concatenate([self.p_u, self.nu + self.p_p])      # dofs.jl:38
        self.inv_p_inversion = invperm(self.p_inversion)


class FEData:
    def __init__(self, mesh: Mesh, spaces: Spaces, dofs: DoFHandler | None = None):
        self.mesh = mesh
        self.spaces = spaces
        self.dofs = dofs if dofs is not None else DoFHandler(*compute_dof_perms(spaces))
