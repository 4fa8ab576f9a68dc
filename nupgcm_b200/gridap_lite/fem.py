"""Minimal affine-simplex Lagrange finite elements (P1 / P2), the part of Gridap that nuPGCM uses
at set-up time.

This is host-side set-up code, not the hot path: it produces the operands (CSR matrices, RHS
vectors, element tables) that the reference obtains from Gridap 0.20.3 / GridapGmsh 0.7.4
(``Manifest.toml:562-584``, not vendored in the reference tree) at
``src/meshes.jl:29-39``, ``src/spaces.jl:31-72``, ``src/inversion.jl:121-249`` and
``src/evolution.jl:209-296``.  The numbering rules follow what was verified against the reference
fixtures:

* vertex ids are gmsh node tags; tetrahedra have their vertex lists sorted ascending (GridapGmsh
  orients 3-D simplices), 2-D cells are left as in the file;
* edges are numbered by first appearance over cells in order and local edges in the order
  (1,2),(1,3),(2,3),(1,4),(2,4),(3,4);
* P2 "owners" are the vertices followed by the edges; vector components are consecutive within
  an owner; Dirichlet DOFs are skipped when numbering free DOFs.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import quadrature
from .mshio import RawMesh

LOCAL_EDGES_3D = np.array([(0, 1), (0, 2), (1, 2), (0, 3), (1, 3), (2, 3)])
LOCAL_EDGES_2D = np.array([(0, 1), (0, 2), (1, 2)])


# --------------------------------------------------------------------------------------------
# topology
# --------------------------------------------------------------------------------------------

def _pair_keys(a, b, nv):
    lo = np.minimum(a, b).astype(np.int64)
    hi = np.maximum(a, b).astype(np.int64)
    return lo * nv + hi


class DiscreteModel:
    """Cells, edges and boundary tags of a simplicial mesh (``GmshDiscreteModel`` stand-in)."""

    def __init__(self, raw: RawMesh, sort_cells: bool | None = None):
        self.dim = raw.dim
        self.nodes = np.ascontiguousarray(raw.nodes, dtype=np.float64)
        cells = raw.elements[self.dim]
        if sort_cells is None:
            sort_cells = self.dim == 3
        if sort_cells:
            cells = np.sort(cells, axis=1)
        self.cells = np.ascontiguousarray(cells, dtype=np.int64)
        self.local_edges = LOCAL_EDGES_3D if self.dim == 3 else LOCAL_EDGES_2D
        self.nv = self.nodes.shape[0]
        self._number_edges()
        self._tag_boundary(raw)
        self._identify_periodic(getattr(raw, "periodic", None))

    # edges by first appearance (cell-major, local-edge-minor)
    def _number_edges(self):
        c = self.cells
        a = c[:, self.local_edges[:, 0]]
        b = c[:, self.local_edges[:, 1]]
        keys = _pair_keys(a, b, self.nv).ravel()
        uniq, first, inv = np.unique(keys, return_index=True, return_inverse=True)
        order = np.argsort(first, kind="stable")
        rank = np.empty_like(order)
        rank[order] = np.arange(order.size)
        self.cell_edges = rank[inv].reshape(a.shape)
        ek = uniq[order]
        self.edges = np.stack([ek // self.nv, ek % self.nv], axis=1)
        self.ne = self.edges.shape[0]
        self._edge_keys_sorted = uniq
        self._edge_ids_sorted = rank

    def edge_ids(self, a, b):
        """Global edge id of vertex pairs (``-1`` when the pair is not an edge)."""
        k = _pair_keys(np.asarray(a), np.asarray(b), self.nv)
        pos = np.searchsorted(self._edge_keys_sorted, k)
        pos = np.minimum(pos, self._edge_keys_sorted.size - 1)
        ok = self._edge_keys_sorted[pos] == k
        return np.where(ok, self._edge_ids_sorted[pos], -1)

    def _tag_boundary(self, raw: RawMesh):
        names = sorted({n for d in range(self.dim) for t in raw.element_names[d] for n in t})
        self.tag_names = names
        bit = {n: np.uint32(1 << i) for i, n in enumerate(names)}
        vtag = np.zeros(self.nv, dtype=np.uint32)
        etag = np.zeros(self.ne, dtype=np.uint32)
        self.boundary = {}
        for d in range(self.dim):
            el = raw.elements[d]
            if not len(el):
                continue
            mask = np.array([np.bitwise_or.reduce([bit[n] for n in t] or [np.uint32(0)])
                             for t in raw.element_names[d]], dtype=np.uint32)
            self.boundary[d] = (el, mask)
            for k in range(d + 1):
                np.bitwise_or.at(vtag, el[:, k], mask)
            for i in range(d + 1):
                for j in range(i + 1, d + 1):
                    eid = self.edge_ids(el[:, i], el[:, j])
                    ok = eid >= 0
                    np.bitwise_or.at(etag, eid[ok], mask[ok])
        self.vertex_tags = vtag
        self.edge_tags = etag
        self._bit = bit

    def _identify_periodic(self, master):
        """Owner-level identification of a periodic mesh: ``vertex_master`` / ``edge_master`` give, for every
        vertex / edge, the vertex / edge whose DOFs it shares (itself when it is not a slave).  An edge is a
        slave when both of its end points are; its master is the edge between their masters, which a
        periodic surface mesh always has."""
        self.vertex_master = np.arange(self.nv, dtype=np.int64)
        self.edge_master = np.arange(self.ne, dtype=np.int64)
        self.periodic = master is not None
        if master is None:
            return
        master = np.asarray(master, dtype=np.int64)
        if master.shape != (self.nv,) or not np.array_equal(master[master], master):
            raise ValueError("periodic: need one master per node, masters being their own masters")
        self.vertex_master = master
        a, b = self.edges[:, 0], self.edges[:, 1]
        slave = (master[a] != a) & (master[b] != b)
        em = self.edge_ids(master[a[slave]], master[b[slave]])
        if (em < 0).any():
            raise ValueError("periodic: a slave edge has no master edge (surface meshes do not match)")
        self.edge_master[slave] = em
        tags_differ = (self.vertex_tags != self.vertex_tags[self.vertex_master]).any() or \
            (self.edge_tags != self.edge_tags[self.edge_master]).any()
        if tags_differ:
            raise ValueError("periodic: slave and master carry different boundary tags")

    def tag_mask(self, tags) -> np.uint32:
        m = np.uint32(0)
        for t in tags:
            if t not in self._bit:
                raise KeyError(f"tag {t!r} not in mesh (have {self.tag_names})")
            m |= self._bit[t]
        return m

    def boundary_facets(self, tags):
        """Facets (dim-1 elements, file vertex order) carrying any of ``tags``."""
        el, mask = self.boundary[self.dim - 1]
        return el[(mask & self.tag_mask(tags)) != 0]

    # geometry --------------------------------------------------------------------------
    def cell_geometry(self, cells=None):
        """Barycentric gradients ``(nc, d+1, 3)`` and measures ``(nc,)`` of affine cells.

        For 2-D cells embedded in 3-D the gradients are the tangential ones."""
        c = self.cells if cells is None else cells
        return simplex_geometry(self.nodes[c])


def simplex_geometry(x):
    """``x``: ``(nc, d+1, 3)`` vertex coordinates -> (grad lambda ``(nc, d+1, 3)``, measure)."""
    nc, nvert, _ = x.shape
    d = nvert - 1
    J = x[:, 1:, :] - x[:, :1, :]                     # (nc, d, 3) edge vectors as rows
    G = np.einsum("cik,cjk->cij", J, J)               # Gram matrix (nc, d, d)
    Ginv = np.linalg.inv(G)
    # gradients of lambda_1..lambda_d: rows of G^{-1} J  (tangential gradient)
    gl = np.einsum("cij,cjk->cik", Ginv, J)           # (nc, d, 3)
    g0 = -gl.sum(axis=1, keepdims=True)
    grad = np.concatenate([g0, gl], axis=1)
    fact = {1: 1.0, 2: 2.0, 3: 6.0}[d]
    meas = np.sqrt(np.abs(np.linalg.det(G))) / fact
    return grad, meas


# --------------------------------------------------------------------------------------------
# reference bases (barycentric form)
# --------------------------------------------------------------------------------------------

def p1_basis(bary):
    """Values ``(nq, d+1)`` and barycentric derivatives ``(nq, d+1, d+1)`` of the P1 basis."""
    nq, nb = bary.shape
    return bary.copy(), np.broadcast_to(np.eye(nb), (nq, nb, nb)).copy()


def p2_basis(bary, local_edges):
    """P2 basis in local order vertices then edges.

    Returns values ``(nq, nloc)`` and derivatives w.r.t. the barycentric coordinates
    ``(nq, nloc, d+1)``; the physical gradient is ``dphi @ grad_lambda``."""
    nq, nb = bary.shape
    ne = len(local_edges)
    val = np.zeros((nq, nb + ne))
    der = np.zeros((nq, nb + ne, nb))
    for i in range(nb):
        val[:, i] = bary[:, i] * (2 * bary[:, i] - 1)
        der[:, i, i] = 4 * bary[:, i] - 1
    for e, (i, j) in enumerate(local_edges):
        val[:, nb + e] = 4 * bary[:, i] * bary[:, j]
        der[:, nb + e, i] = 4 * bary[:, j]
        der[:, nb + e, j] = 4 * bary[:, i]
    return val, der


# --------------------------------------------------------------------------------------------
# spaces
# --------------------------------------------------------------------------------------------

class LagrangeSpace:
    """P1/P2 Lagrange space with ``ncomp`` components and per-component Dirichlet masks.

    ``owner_dofs[o, c]`` is the free DOF id (``>= 0``) or ``-(k+1)`` for Dirichlet DOF ``k``
    (Gridap's convention of negative ids for Dirichlet DOFs).
    """

    def __init__(self, model: DiscreteModel, order: int, ncomp: int = 1,
                 dirichlet_tags=(), dirichlet_masks=None, fix_last_owner: bool = False):
        self.model = model
        self.order = order
        self.ncomp = ncomp
        nv, ne = model.nv, model.ne
        if order == 1:
            self.n_owners = nv
            self.cell_owners = model.cells
            owner_tags = model.vertex_tags
            owner_master = model.vertex_master
        elif order == 2:
            self.n_owners = nv + ne
            self.cell_owners = np.concatenate([model.cells, nv + model.cell_edges], axis=1)
            owner_tags = np.concatenate([model.vertex_tags, model.edge_tags])
            owner_master = np.concatenate([model.vertex_master, nv + model.edge_master])
        else:
            raise ValueError("order must be 1 or 2")
        diri = np.zeros((self.n_owners, ncomp), dtype=bool)
        tags = list(dirichlet_tags)
        if dirichlet_masks is None:
            dirichlet_masks = [(True,) * ncomp] * len(tags)
        for tag, mask in zip(tags, dirichlet_masks):
            on = (owner_tags & model.tag_mask([tag])) != 0
            for c in range(ncomp):
                if mask[c]:
                    diri[on, c] = True
        if fix_last_owner:           # constraint=:zeromean fixes the last DOF (spaces.jl:45)
            diri[owner_master[-1], :] = True
        # periodic meshes: a slave owner has no DOFs of its own, it carries its master's ids
        own = owner_master == np.arange(self.n_owners)
        self.owner_is_master = own
        flat = diri.ravel()
        numbered = np.repeat(own, ncomp)
        ids = np.zeros(flat.size, dtype=np.int64)
        free_sel, diri_sel = numbered & ~flat, numbered & flat
        ids[free_sel] = np.arange(free_sel.sum())
        ids[diri_sel] = -(np.arange(diri_sel.sum()) + 1)
        ids = ids.reshape(self.n_owners, ncomp)
        self.owner_dofs = ids[owner_master]
        self.nfree = int(free_sel.sum())
        self.ndiri = int(diri_sel.sum())
        self.dirichlet_values = np.zeros(self.ndiri)

    def owner_coordinates(self):
        m = self.model
        if self.order == 1:
            return m.nodes
        mid = 0.5 * (m.nodes[m.edges[:, 0]] + m.nodes[m.edges[:, 1]])
        return np.concatenate([m.nodes, mid], axis=0)

    def _owner_tags(self):
        m = self.model
        return m.vertex_tags if self.order == 1 else np.concatenate([m.vertex_tags, m.edge_tags])

    def _evaluate(self, fun, x):
        """``fun``: callable of ``(n, 3)`` points or a constant; result broadcast to (n, ncomp)."""
        v = np.asarray(fun(x) if callable(fun) else fun, dtype=np.float64)
        if v.ndim == 1 and v.shape[0] == x.shape[0] and not (v.shape[0] == self.ncomp > 1):
            v = v.reshape(-1, 1)
        return np.broadcast_to(v, (x.shape[0], self.ncomp))

    def interpolate(self, fun):
        """Nodal interpolation; returns the (free, Dirichlet) value vectors."""
        v = self._evaluate(fun, self.owner_coordinates())
        free = np.zeros(self.nfree)
        dval = np.zeros(self.ndiri)
        ids = self.owner_dofs
        free[ids[ids >= 0]] = v[ids >= 0]
        dval[-ids[ids < 0] - 1] = v[ids < 0]
        return free, dval

    def set_dirichlet(self, tags, funs):
        """Dirichlet values tag by tag (``TrialFESpace(space, values)``, spaces.jl:54-64)."""
        x = self.owner_coordinates()
        owner_tags = self._owner_tags()
        for tag, f in zip(tags, funs):
            on = (owner_tags & self.model.tag_mask([tag])) != 0
            val = self._evaluate(f, x[on])
            ids = self.owner_dofs[on]
            sel = ids < 0
            self.dirichlet_values[-ids[sel] - 1] = val[sel]

    def full_values(self, free, dirichlet=None):
        """Owner-major array ``(n_owners, ncomp)`` of all DOF values (free + Dirichlet)."""
        dval = self.dirichlet_values if dirichlet is None else dirichlet
        ids = self.owner_dofs
        out = np.empty(ids.shape)
        out[ids >= 0] = np.asarray(free)[ids[ids >= 0]]
        out[ids < 0] = dval[-ids[ids < 0] - 1]
        return out


# --------------------------------------------------------------------------------------------
# owner-level assembly
# --------------------------------------------------------------------------------------------

def _chunks(n, size):
    for s in range(0, n, size):
        yield slice(s, min(n, s + size))


class CellIntegrator:
    """Quadrature data shared by every form on one set of affine cells."""

    def __init__(self, model: DiscreteModel, tet_rule: str | None = None, chunk: int = 16384):
        self.model = model
        self.bary, self.w = quadrature.simplex(model.dim, tet_rule)
        self.rule = tet_rule or ("keast11" if model.dim == 3 else "tri6")
        self.chunk = chunk
        self.phi2, self.dphi2 = p2_basis(self.bary, model.local_edges)
        self.phi1, self.dphi1 = p1_basis(self.bary)
        self.grad, self.meas = model.cell_geometry()

    def basis(self, order):
        return (self.phi1, self.dphi1) if order == 1 else (self.phi2, self.dphi2)

    def xq(self, sl):
        """Quadrature point coordinates ``(nc, nq, 3)`` of a cell slice."""
        return np.einsum("qk,ckd->cqd", self.bary, self.model.nodes[self.model.cells[sl]])

    def coefficient(self, c, sl):
        """Evaluate a coefficient (callable of ``(n,3)`` points, or scalar) at quadrature points."""
        nc = self.model.cells[sl].shape[0]
        if callable(c):
            x = self.xq(sl)
            return np.asarray(c(x.reshape(-1, 3)), dtype=np.float64).reshape(nc, -1)
        return np.full((nc, self.w.size), float(c))

    def matrix(self, kind, test: LagrangeSpace, trial: LagrangeSpace, coef=1.0, dirs=(0, 1, 2),
               comp=None):
        """Owner-level sparse matrix ``(test.n_owners, trial.n_owners)`` keeping every
        (test, trial) owner pair that shares a cell, numerically zero or not.

        kind: ``"mass"`` c φᵢφⱼ ; ``"stiff"`` c Σ_{d∈dirs} ∂dφᵢ ∂dφⱼ ;
        ``"grad_test"`` c (∂_comp φᵢ) φⱼ ; ``"grad_trial"`` c φᵢ (∂_comp φⱼ) ;
        ``"dd"`` c ∂_{comp[0]}φᵢ ∂_{comp[1]}φⱼ.
        """
        vt, dt = self.basis(test.order)
        vr, dr = self.basis(trial.order)
        rows, cols, vals = [], [], []
        dirs = list(dirs)
        for sl in _chunks(self.model.cells.shape[0], self.chunk):
            wq = self.coefficient(coef, sl) * self.w[None, :] * self.meas[sl, None]  # (nc, nq)
            g = self.grad[sl]
            if kind == "mass":
                ke = np.einsum("cq,qi,qj->cij", wq, vt, vr)
            elif kind == "stiff":
                gt = np.einsum("qik,ckd->cqid", dt, g[:, :, dirs])
                gr = gt if trial.order == test.order else np.einsum("qik,ckd->cqid", dr,
                                                                    g[:, :, dirs])
                ke = np.einsum("cq,cqid,cqjd->cij", wq, gt, gr)
            elif kind == "dd":            # c ∂_{comp[0]} φᵢ ∂_{comp[1]} φⱼ
                gt = np.einsum("qik,ck->cqi", dt, g[:, :, comp[0]])
                gr = np.einsum("qjk,ck->cqj", dr, g[:, :, comp[1]])
                ke = np.einsum("cq,cqi,cqj->cij", wq, gt, gr)
            elif kind == "grad_test":
                gt = np.einsum("qik,ck->cqi", dt, g[:, :, comp])
                ke = np.einsum("cq,cqi,qj->cij", wq, gt, vr)
            elif kind == "grad_trial":
                gr = np.einsum("qjk,ck->cqj", dr, g[:, :, comp])
                ke = np.einsum("cq,qi,cqj->cij", wq, vt, gr)
            else:
                raise ValueError(kind)
            ct = test.cell_owners[sl]
            cr = trial.cell_owners[sl]
            rows.append(np.repeat(ct, cr.shape[1], axis=1).ravel())
            cols.append(np.tile(cr, (1, ct.shape[1])).ravel())
            vals.append(ke.ravel())
        m = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                          shape=(test.n_owners, trial.n_owners))
        out = m.tocsr()            # sums duplicates, keeps explicit zeros
        out.sort_indices()
        return out

    def vector(self, kind, test: LagrangeSpace, coef=1.0, comp=None):
        """Owner-level vector: ``"value"`` ∫ c φᵢ ; ``"grad"`` ∫ c ∂_comp φᵢ."""
        vt, dt = self.basis(test.order)
        out = np.zeros(test.n_owners)
        for sl in _chunks(self.model.cells.shape[0], self.chunk):
            wq = self.coefficient(coef, sl) * self.w[None, :] * self.meas[sl, None]
            if kind == "value":
                fe = np.einsum("cq,qi->ci", wq, vt)
            elif kind == "grad":
                gt = np.einsum("qik,ck->cqi", dt, self.grad[sl][:, :, comp])
                fe = np.einsum("cq,cqi->ci", wq, gt)
            else:
                raise ValueError(kind)
            np.add.at(out, test.cell_owners[sl].ravel(), fe.ravel())
        return out


class FacetIntegrator:
    """``∫_Γ g φᵢ dΓ`` over tagged boundary facets (``BoundaryTriangulation(model; tags)``)."""

    def __init__(self, model: DiscreteModel, tags):
        self.model = model
        self.facets = model.boundary_facets(tags)
        d = model.dim - 1
        self.bary, self.w = quadrature.simplex(d)
        self.local_edges = LOCAL_EDGES_2D if d == 2 else np.array([(0, 1)])
        x = model.nodes[self.facets]
        _, self.meas = simplex_geometry(x) if len(x) else (None, np.zeros(0))
        self.phi2, _ = p2_basis(self.bary, self.local_edges)

    def facet_owners(self, order):
        f = self.facets
        if order == 1:
            return f
        e = [self.model.edge_ids(f[:, i], f[:, j]) for i, j in self.local_edges]
        return np.concatenate([f, self.model.nv + np.stack(e, axis=1)], axis=1)

    def vector(self, test: LagrangeSpace, g):
        out = np.zeros(test.n_owners)
        if not len(self.facets):
            return out
        xq = np.einsum("qk,fkd->fqd", self.bary, self.model.nodes[self.facets])
        gq = np.asarray(g(xq.reshape(-1, 3)), dtype=np.float64)
        gq = np.broadcast_to(gq, (xq.shape[0] * xq.shape[1],)).reshape(xq.shape[:2])
        phi = self.phi2 if test.order == 2 else self.bary
        fe = np.einsum("fq,q,f,qi->fi", gq, self.w, self.meas, phi)
        np.add.at(out, self.facet_owners(test.order).ravel(), fe.ravel())
        return out


def restrict(mat_owner, test: LagrangeSpace, trial: LagrangeSpace, blocks):
    """Expand an owner-level pattern to component DOFs and split free / Dirichlet columns.

    ``blocks[(a, b)]`` is an owner-level CSR matrix (all with the pattern of ``mat_owner``) giving
    the coupling of test component ``a`` with trial component ``b``; missing pairs are stored
    as explicit zeros, which is what Gridap's symbolic assembly does (SURVEY.md finding 8).
    Returns ``(A_ff, A_fd)`` in free-DOF numbering (``A_fd`` has one column per Dirichlet DOF).
    """
    pat = mat_owner.tocoo()
    nnz = pat.nnz
    nt, nr = test.ncomp, trial.ncomp
    data = np.zeros((nnz, nt, nr))
    for (a, b), m in blocks.items():
        m = m.tocsr()
        m.sort_indices()
        if m.nnz != nnz:
            raise ValueError("component blocks must share the owner-level pattern")
        data[:, a, b] = m.data
    r = test.owner_dofs[pat.row][:, :, None].repeat(nr, axis=2)      # (nnz, nt, nr)
    c = trial.owner_dofs[pat.col][:, None, :].repeat(nt, axis=1)
    r, c, v = r.ravel(), c.ravel(), data.ravel()
    rf = r >= 0
    ff = rf & (c >= 0)
    fd = rf & (c < 0)
    a_ff = sp.coo_matrix((v[ff], (r[ff], c[ff])), shape=(test.nfree, trial.nfree)).tocsr()
    a_fd = sp.coo_matrix((v[fd], (r[fd], -c[fd] - 1)), shape=(test.nfree, trial.ndiri)).tocsr()
    a_ff.sort_indices()
    return a_ff, a_fd


def restrict_vector(vec_owner_comp, test: LagrangeSpace):
    """``(n_owners, ncomp)`` owner-level vector -> free-DOF vector."""
    v = np.asarray(vec_owner_comp).reshape(test.n_owners, test.ncomp)
    out = np.zeros(test.nfree)
    ids = test.owner_dofs
    if test.owner_is_master.all():
        out[ids[ids >= 0]] = v[ids >= 0]
    else:                            # periodic: the slave owners' integrals belong to their masters' DOFs
        np.add.at(out, ids[ids >= 0], v[ids >= 0])
    return out
