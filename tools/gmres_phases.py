import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nupgcm_b200 import lib, workloads as W
from nupgcm_b200.architectures import GPU
ctx = GPU(0).ctx
w = W.bowl_example(h=0.08); ops = W.host_operands(w)
A = ops["A"]; y = ops["B"] @ ops["b_init"] + ops["b0"]
for drop in (True,):
    dA = ctx.csr(A, drop_zeros=drop); dy = ctx.vector(y)
    for grid in (148,):
        os.environ["NUPGCM_GRID"] = str(grid)
        for orth, name in ((lib.ORTH_MGS, "mgs"), (lib.ORTH_CGS2, "cgs2"), (lib.ORTH_CGS2_FUSED, "cgs2f")):
            x = ctx.vector(y.size)
            lib.gmres_solve(dA, dy, x, pscale=ops["pscale"], atol=0, rtol=1e-30, itmax=100, orth=orth)
            x = ctx.vector(y.size)
            st, _ = lib.gmres_solve(dA, dy, x, pscale=ops["pscale"], atol=0, rtol=1e-30, itmax=1000, orth=orth)
            us = 1e3 * st.device_ms / st.niter
            pf = list(st.phase_frac)
            print(f"drop={int(drop)} grid={grid} {name:4s}: {us:6.2f} us/iter  spmv {us*pf[0]:5.2f}  local {us*pf[1]:5.2f}  reduce {us*pf[2]:5.2f}  scalar {us*pf[3]:5.2f}  SM {st.sm_mhz:.0f} MHz", flush=True)
