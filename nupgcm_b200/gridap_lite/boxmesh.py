"""Structured tetrahedral box mesh with the boundary names of the reference's meshes.

BASELINE config 4 asks for the ``channel_basin`` mesh (reference ``meshes/channel_basin_flat.jl``:
x ∈ [0, 1], y ∈ [−1, 1], flat bottom at z = −α, generated with gmsh).  gmsh is not available, so
the declared substitution is a structured box of the same extent: every grid cell
is cut into six tetrahedra around its main diagonal (Kuhn triangulation, conforming across cells),
the top face carries the physical name ``surface``, floor and walls ``bottom``, and the rim where
they meet is the ``coastline`` curve — the three names the reference's ``Spaces`` keys its
Dirichlet conditions on (``scratch/run.jl:132-136``).

``periodic_x_below=y0`` reproduces the channel of the reference mesh (``meshes/channel_basin_flat.jl:114-121``:
``setPeriodic`` of the east wall onto the west wall for y <= y0 = -L/4): the wall faces with y <= y0 are no
boundary (the reference tags them "interior"), their rim at the surface is no coastline, and every node of
the east face with y <= y0 is the slave of the west-face node at the same (y, z) — ``RawMesh.periodic``.  The
Kuhn triangulation is translation invariant, so the two surface meshes match as gmsh's periodic meshes do.
"""
from __future__ import annotations

import itertools

import numpy as np

from .mshio import RawMesh


def box_mesh(nx: int, ny: int, nz: int, x=(0.0, 1.0), y=(-1.0, 1.0), z=(-0.125, 0.0),
             periodic_x_below: float | None = None) -> RawMesh:
    if periodic_x_below is not None and nx < 3:
        raise ValueError("a periodic direction needs at least 3 cells (distinct edges after identification)")
    xs = np.linspace(x[0], x[1], nx + 1)
    ys = np.linspace(y[0], y[1], ny + 1)
    zs = np.linspace(z[0], z[1], nz + 1)
    X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij")
    nodes = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)

    def nid(i, j, k):
        return (i * (ny + 1) + j) * (nz + 1) + k

    I, J, K = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    I, J, K = I.ravel(), J.ravel(), K.ravel()
    tets = []
    for perm in itertools.permutations(range(3)):           # six paths 000 -> 111
        off = np.zeros(3, dtype=np.int64)
        verts = [nid(I, J, K)]
        for ax in perm:
            off[ax] = 1
            verts.append(nid(I + off[0], J + off[1], K + off[2]))
        tets.append(np.stack(verts, axis=1))
    tets = np.concatenate(tets, axis=0).astype(np.int64)
    # positive orientation is irrelevant to the FE code (cells are sorted by node id)

    # boundary faces: faces that belong to exactly one tetrahedron
    faces = np.concatenate([tets[:, [1, 2, 3]], tets[:, [0, 2, 3]], tets[:, [0, 1, 3]], tets[:, [0, 1, 2]]])
    key = np.sort(faces, axis=1)
    uniq, first, counts = np.unique(key, axis=0, return_index=True, return_counts=True)
    bnd = faces[np.sort(first[counts == 1])]
    tol = 1e-12 * max(abs(y[0]), abs(y[1]), 1.0)
    in_channel = (lambda pts: np.all(pts[..., 1] <= periodic_x_below + tol, axis=-1)) if periodic_x_below is not None \
        else (lambda pts: np.zeros(pts.shape[:-2], dtype=bool))
    on_x_wall = lambda pts: (np.all(np.isclose(pts[..., 0], x[0]), axis=-1) |                 # noqa: E731
                             np.all(np.isclose(pts[..., 0], x[1]), axis=-1))
    bnd = bnd[~(on_x_wall(nodes[bnd]) & in_channel(nodes[bnd]))]        # periodic faces are interior
    on_top = np.all(np.isclose(nodes[bnd][:, :, 2], z[1]), axis=1)
    tri_names = [("surface",) if t else ("bottom",) for t in on_top]

    # coastline: edges of surface triangles lying on the rim of the top face
    top = bnd[on_top]
    e = np.concatenate([top[:, [0, 1]], top[:, [0, 2]], top[:, [1, 2]]])
    p, q = nodes[e[:, 0]], nodes[e[:, 1]]
    rim = np.zeros(len(e), dtype=bool)
    for ax, lim in ((0, x), (1, y)):
        for v in lim:
            rim |= np.isclose(p[:, ax], v) & np.isclose(q[:, ax], v)
    pq = np.stack([p, q], axis=1)
    rim &= ~(on_x_wall(pq) & in_channel(pq))                            # the channel's side rims are open
    lines = np.unique(np.sort(e[rim], axis=1), axis=0)
    line_names = [("coastline",)] * len(lines)
    pts = np.unique(lines.ravel())[:, None]
    pt_names = [("coastline",)] * len(pts)
    periodic = None
    if periodic_x_below is not None:
        periodic = np.arange(len(nodes), dtype=np.int64)
        jj, kk = np.meshgrid(np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
        jj, kk = jj.ravel(), kk.ravel()
        sel = ys[jj] <= periodic_x_below + tol
        periodic[nid(nx, jj[sel], kk[sel])] = nid(0, jj[sel], kk[sel])
    return RawMesh(periodic=periodic, nodes=nodes,
                   elements={0: pts.astype(np.int64), 1: lines.astype(np.int64), 2: bnd.astype(np.int64), 3: tets},
                   element_names={0: pt_names, 1: line_names, 2: tri_names,
                                  3: [("interior",)] * len(tets)},
                   physical_names={(0, 1): "coastline", (1, 1): "coastline", (2, 1): "surface",
                                   (2, 2): "bottom", (3, 1): "interior"})
