"""GMRES(20) per-iteration time on the h = 0.04 inversion system against the host-side switches that change the
load of the grid reductions: CTAs per GPU (NUPGCM_GRID), replicas of the reduction slots (NUPGCM_REPLICAS) and the
back-off between two polls of a flagged word (NUPGCM_POLL_SLEEP, ns).  All three are read per solve.

    python tools/reduce_sweep.py [level] [iters] [grids, comma separated] [replicas] [sleeps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nupgcm_b200 import lib, workloads as W          # noqa: E402
from nupgcm_b200.architectures import GPU            # noqa: E402
from nupgcm_b200.inversion import permuted_inversion_system   # noqa: E402

level = int(sys.argv[1]) if len(sys.argv) > 1 else 1
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 400
ctx = GPU(0).ctx
w = W.bowl_example(mesh=W.refined_bowl(level)) if level > 0 else W.bowl_example()
fe = w.fe_data()
A, B, b0, pscale = permuted_inversion_system(fe, w.params, w.forcings)
y = B @ fe.spaces.B.interpolate(w.b0)[0][fe.dofs.p_b] + b0
dA = ctx.csr(A, drop_zeros=True)
dy = ctx.vector(y)
x = ctx.vector(y.size)
print(f"h = {0.08 / 2 ** level:g}: N = {y.size}, {iters} iterations per solve; us per iteration", flush=True)


def run(orth):
    lib.gmres_solve(dA, dy, x, pscale=pscale, atol=0, rtol=1e-30, itmax=40, orth=orth)
    st, _ = lib.gmres_solve(dA, dy, x, pscale=pscale, atol=0, rtol=1e-30, itmax=iters, orth=orth)
    return 1e3 * st.device_ms / st.niter, st.rnorm / st.rnorm0


_list = lambda i, d: tuple(int(v) for v in sys.argv[i].split(",")) if len(sys.argv) > i else d      # noqa: E731
for grid in _list(3, (148, 140, 128, 112, 96)):
    os.environ["NUPGCM_GRID"] = str(grid)
    for rep in _list(4, (1, 2, 4)):
        os.environ["NUPGCM_REPLICAS"] = str(rep)
        row = []
        for sleep in _list(5, (0, 100, 300, 600)):
            os.environ["NUPGCM_POLL_SLEEP"] = str(sleep)
            t, r = run(lib.ORTH_MGS)
            row.append(f"sleep {sleep:3d}: {t:6.1f}")
        os.environ["NUPGCM_POLL_SLEEP"] = "0"
        tc, _ = run(lib.ORTH_CGS2_FUSED)
        print(f"grid {grid:3d} replicas {rep}: mgs " + "  ".join(row) + f"   | cgs2f (spin) {tc:6.1f}   rnorm/rnorm0 {r:.3e}", flush=True)
