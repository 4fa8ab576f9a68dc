"""MSH 4.1 (ASCII) reader.

Stands in for ``GmshDiscreteModel(ifile)`` (reference ``src/meshes.jl:30``): the reference
hands the ``.msh`` file to GridapGmsh, which is not available here, so the file is parsed
directly.  Only what the nuPGCM meshes use is supported: points, lines, triangles and
tetrahedra classified on entities that carry physical names.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# gmsh element type -> (topological dimension, nodes per element)
_ELEM = {15: (0, 1), 1: (1, 2), 2: (2, 3), 4: (3, 4)}


@dataclass
class RawMesh:
    """Plain arrays of a gmsh mesh.

    ``elements[d]`` is an ``(n, d+1)`` int64 array of 0-based node ids of the ``d``-dimensional
    elements, in file order; ``element_names[d]`` holds, per element, the tuple of physical names
    of the entity the element is classified on.
    """

    nodes: np.ndarray
    elements: dict = field(default_factory=dict)
    element_names: dict = field(default_factory=dict)
    physical_names: dict = field(default_factory=dict)  # (dim, tag) -> name
    # periodic meshes (gmsh.model.mesh.setPeriodic, reference meshes/channel_basin_flat.jl:114-121): master node of
    # every node (itself when it is not a slave); None = not periodic.  Geometry keeps all nodes, the FE
    # spaces give a slave node (and a slave edge) the DOFs of its master.
    periodic: np.ndarray | None = None

    @property
    def dim(self) -> int:
        return max(d for d, e in self.elements.items() if len(e))

    # compact binary form (the repository ships meshes as .npz, see tools/make_fixtures.py)
    def save_npz(self, path: str):
        names = sorted({n for d in range(4) for t in self.element_names.get(d, []) for n in t})
        bit = {n: 1 << i for i, n in enumerate(names)}
        out = {"nodes": self.nodes, "names": np.array(names)}
        for d in range(4):
            el = self.elements.get(d, np.zeros((0, d + 1), dtype=np.int64))
            out[f"elements{d}"] = el.astype(np.int32)
            out[f"mask{d}"] = np.array([sum(bit[n] for n in t) for t in self.element_names.get(d, [])],
                                       dtype=np.uint32)
        np.savez_compressed(path, **out)

    @staticmethod
    def load_npz(path: str) -> "RawMesh":
        z = np.load(path)
        names = [str(n) for n in z["names"]]
        elements, element_names = {}, {}
        for d in range(4):
            elements[d] = z[f"elements{d}"].astype(np.int64)
            element_names[d] = [tuple(n for i, n in enumerate(names) if m >> i & 1)
                                for m in z[f"mask{d}"].tolist()]
        return RawMesh(nodes=z["nodes"], elements=elements, element_names=element_names)


def _sections(text: str) -> dict:
    out = {}
    lines = text.splitlines()
    i = 0
    while i < len(lines):
        ln = lines[i].strip()
        if ln.startswith("$") and not ln.startswith("$End"):
            name = ln[1:]
            j = i + 1
            while lines[j].strip() != "$End" + name:
                j += 1
            out[name] = lines[i + 1:j]
            i = j
        i += 1
    return out


def read_msh(path: str) -> RawMesh:
    with open(path, "r") as fh:
        sec = _sections(fh.read())
    version = sec["MeshFormat"][0].split()[0]
    if not version.startswith("4.1"):
        raise ValueError(f"{path}: only MSH 4.1 ASCII is supported (found {version})")

    phys = {}
    for ln in sec.get("PhysicalNames", [])[1:]:
        d, tag, name = ln.split(maxsplit=2)
        phys[(int(d), int(tag))] = name.strip().strip('"')

    # entity (dim, tag) -> physical tags
    ent_phys = {}
    ent = sec["Entities"]
    counts = [int(v) for v in ent[0].split()]
    row = 1
    for d, cnt in enumerate(counts):
        for _ in range(cnt):
            tok = ent[row].split()
            row += 1
            tag = int(tok[0])
            k = 4 if d == 0 else 7          # points carry x y z, the rest a bounding box
            nphys = int(tok[k])
            ent_phys[(d, tag)] = [abs(int(v)) for v in tok[k + 1:k + 1 + nphys]]

    nod = sec["Nodes"]
    nblocks, nnodes, _, maxtag = (int(v) for v in nod[0].split())
    if maxtag != nnodes:
        raise ValueError("non-contiguous node tags are not supported")
    xyz = np.zeros((nnodes, 3))
    row = 1
    for _ in range(nblocks):
        _, _, parametric, n = (int(v) for v in nod[row].split())
        if parametric:
            raise ValueError("parametric node blocks are not supported")
        tags = np.array([int(v) for v in nod[row + 1:row + 1 + n]], dtype=np.int64)
        if n:
            xyz[tags - 1] = np.array([[float(v) for v in ln.split()[:3]]
                                      for ln in nod[row + 1 + n:row + 1 + 2 * n]])
        row += 1 + 2 * n

    elems = {d: [] for d in range(4)}
    names = {d: [] for d in range(4)}
    ele = sec["Elements"]
    nblocks = int(ele[0].split()[0])
    row = 1
    for _ in range(nblocks):
        edim, etag, etype, n = (int(v) for v in ele[row].split())
        if etype not in _ELEM:
            raise ValueError(f"unsupported gmsh element type {etype}")
        d, nn = _ELEM[etype]
        block_names = tuple(phys[(edim, t)] for t in ent_phys.get((edim, etag), [])
                            if (edim, t) in phys)
        for ln in ele[row + 1:row + 1 + n]:
            tok = ln.split()
            elems[d].append([int(v) - 1 for v in tok[1:1 + nn]])
            names[d].append(block_names)
        row += 1 + n

    elements = {d: np.array(v, dtype=np.int64).reshape(-1, d + 1) for d, v in elems.items()}
    return RawMesh(nodes=xyz, elements=elements, element_names=names, physical_names=phys)
