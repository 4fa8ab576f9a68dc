"""Uniform (red) refinement of simplicial meshes.

The reference generates its finer meshes with gmsh (``meshes/mesh_bowl3D.jl:1-47``), which is not
available here; BASELINE configs 3 and 5 (h = 0.04, h = 0.02) are obtained instead by refining the
shipped h = 0.08 mesh once or twice.  Every tetrahedron is split into 8 (four
corner tets plus the inner octahedron cut along its m02-m13 diagonal), every boundary triangle
into 4, every line into 2; physical names are inherited, old nodes keep their ids and the edge
midpoints are appended in order of first appearance.  ``project(x, names)`` may move the new
boundary nodes onto the true geometry.
"""
from __future__ import annotations

import numpy as np

from .mshio import RawMesh

_TET_EDGES = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
# children in terms of local ids 0-3 (vertices) and 4-9 (midpoints of _TET_EDGES in that order)
_TET_CHILDREN = [(0, 4, 5, 6), (4, 1, 7, 8), (5, 7, 2, 9), (6, 8, 9, 3),
                 (4, 5, 6, 8), (4, 5, 7, 8), (5, 6, 8, 9), (5, 7, 8, 9)]
_TRI_EDGES = [(0, 1), (0, 2), (1, 2)]
_TRI_CHILDREN = [(0, 3, 4), (3, 1, 5), (4, 5, 2), (3, 4, 5)]


def _edge_midpoint_ids(elems_by_dim, n0):
    """Global midpoint node id of every edge, numbered by first appearance (highest dim first)."""
    keys = []
    for d, edges in ((3, _TET_EDGES), (2, _TRI_EDGES), (1, [(0, 1)])):
        el = elems_by_dim.get(d)
        if el is None or not len(el):
            continue
        for i, j in edges:
            a, b = el[:, i], el[:, j]
            keys.append(np.minimum(a, b) * np.int64(n0) + np.maximum(a, b))
    allk = np.concatenate(keys)
    # first-appearance order: element-major within each dimension block is not required, any
    # deterministic order works because midpoint ids are only labels
    uniq, first = np.unique(allk, return_index=True)
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    return uniq, rank


def refine(raw: RawMesh, project=None) -> RawMesh:
    n0 = raw.nodes.shape[0]
    elems = {d: raw.elements.get(d, np.zeros((0, d + 1), dtype=np.int64)) for d in range(4)}
    uniq, rank = _edge_midpoint_ids(elems, n0)

    def mid(a, b):
        k = np.minimum(a, b) * np.int64(n0) + np.maximum(a, b)
        return n0 + rank[np.searchsorted(uniq, k)]

    ek = uniq[np.argsort(rank)]
    ea, eb = ek // n0, ek % n0
    new_nodes = 0.5 * (raw.nodes[ea] + raw.nodes[eb])
    nodes = np.concatenate([raw.nodes, new_nodes], axis=0)

    out_el, out_names = {}, {}
    # points
    out_el[0], out_names[0] = elems[0].copy(), list(raw.element_names.get(0, []))
    # lines
    ln = elems[1]
    if len(ln):
        m = mid(ln[:, 0], ln[:, 1])
        out_el[1] = np.stack([np.stack([ln[:, 0], m], 1), np.stack([m, ln[:, 1]], 1)], axis=1).reshape(-1, 2)
        out_names[1] = [t for t in raw.element_names[1] for _ in range(2)]
    else:
        out_el[1], out_names[1] = ln, []
    # triangles
    tr = elems[2]
    if len(tr):
        loc = np.concatenate([tr] + [mid(tr[:, i], tr[:, j])[:, None] for i, j in _TRI_EDGES], axis=1)
        out_el[2] = loc[:, np.array(_TRI_CHILDREN)].reshape(-1, 3)
        out_names[2] = [t for t in raw.element_names[2] for _ in range(4)]
    else:
        out_el[2], out_names[2] = tr, []
    # tetrahedra
    te = elems[3]
    if len(te):
        loc = np.concatenate([te] + [mid(te[:, i], te[:, j])[:, None] for i, j in _TET_EDGES], axis=1)
        out_el[3] = loc[:, np.array(_TET_CHILDREN)].reshape(-1, 4)
        out_names[3] = [t for t in raw.element_names[3] for _ in range(8)]
    else:
        out_el[3], out_names[3] = te, []

    if project is not None:
        # names carried by each new node: union over the boundary elements whose edge it bisects
        tags = [set() for _ in range(nodes.shape[0] - n0)]
        for d, edges in ((1, [(0, 1)]), (2, _TRI_EDGES)):
            el = elems[d]
            for i, j in edges:
                if not len(el):
                    continue
                ids = mid(el[:, i], el[:, j]) - n0
                for e, names in zip(ids.tolist(), raw.element_names[d]):
                    tags[e].update(names)
        nodes[n0:] = project(nodes[n0:], tags)
    return RawMesh(nodes=nodes, elements={d: np.ascontiguousarray(v, dtype=np.int64) for d, v in out_el.items()},
                   element_names=out_names, physical_names=dict(raw.physical_names))


def bowl_projection(α: float):
    """Move new boundary nodes of the bowl onto z = −α(1 − x² − y²) (bottom) and the unit circle
    (coastline), the geometry of ``meshes/mesh_bowl3D.jl:12-27``."""
    def project(x, tags):
        x = x.copy()
        for i, t in enumerate(tags):
            if "coastline" in t:
                r = np.hypot(x[i, 0], x[i, 1])
                x[i, 0] /= r
                x[i, 1] /= r
                x[i, 2] = 0.0
            elif "bottom" in t:
                x[i, 2] = -α * (1.0 - x[i, 0] ** 2 - x[i, 1] ** 2)
        return x
    return project
