#!/usr/bin/env python
"""Turn the reference's data files into the fixtures this repository ships.

Run in the build container (needs /root/reference, read-only):

    python tools/make_fixtures.py

* meshes/*.msh (gmsh output, git-ignored upstream but shipped in the reference checkout)
    -> nupgcm_b200/data/meshes/*.npz      (inputs of the workloads; parsed by gridap_lite.mshio)
* test/data/*.jld2 (the reference's regression fixtures, decoded by oracle/jld2.py)
    -> tests/golden/*.npz                  (golden vectors of the parity tests)

Nothing under /root/reference is executed (there is no Julia here); the files are only decoded.
"""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nupgcm_b200.gridap_lite import read_msh          # noqa: E402
from oracle.jld2 import read_jld2                      # noqa: E402

REF = "/root/reference"
MESHES = {
    "bowl2D_1.000000e-01_5.000000e-01.msh": "bowl2D_h0.10.npz",
    "bowl3D_1.000000e-01_5.000000e-01.msh": "bowl3D_h0.10.npz",
    "bowl3D_8.000000e-02_5.000000e-01.msh": "bowl3D_h0.08.npz",
}


def main():
    out_m = os.path.join(ROOT, "nupgcm_b200", "data", "meshes")
    out_g = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_m, exist_ok=True)
    os.makedirs(out_g, exist_ok=True)
    for src, dst in MESHES.items():
        raw = read_msh(os.path.join(REF, "meshes", src))
        raw.save_npz(os.path.join(out_m, dst))
        print(f"{src} -> {dst}: {raw.nodes.shape[0]} nodes, "
              f"{raw.elements[raw.dim].shape[0]} cells")
    for f in sorted(glob.glob(os.path.join(REF, "test", "data", "*.jld2"))):
        d = read_jld2(f)
        flat = {}
        for k, v in d.items():
            if isinstance(v, tuple) and v[0] == "csc":
                _, m, n, colptr, rowval, nzval = v
                flat[k + "_shape"] = np.array([m, n])
                flat[k + "_colptr"] = colptr - 1        # 0-based
                flat[k + "_rowval"] = rowval - 1
                flat[k + "_nzval"] = nzval
            elif k == "iperm":
                flat[k] = np.asarray(v) - 1             # 0-based
            else:
                flat[k] = np.asarray(v)
        name = os.path.basename(f).replace(".jld2", ".npz")
        np.savez_compressed(os.path.join(out_g, name), **flat)
        print(f"{os.path.basename(f)} -> tests/golden/{name}: "
              + ", ".join(f"{k}{np.shape(v)}" for k, v in flat.items()))


if __name__ == "__main__":
    main()
