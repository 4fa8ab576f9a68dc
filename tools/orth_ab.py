#!/usr/bin/env python
"""µs per GMRES(20) iteration for the three orthogonalisation variants (MGS / CGS2 / fused CGS2)
on the bowl3D h=0.08 inversion (SM-resident) and, optionally, the refined h=0.04 one (streaming)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nupgcm_b200 import lib, workloads as W            # noqa: E402
from nupgcm_b200.architectures import GPU              # noqa: E402


def main():
    levels = [int(v) for v in sys.argv[1:]] or [0]
    ctx = GPU(0).ctx
    for lv in levels:
        w = W.bowl_example(h=0.08) if lv == 0 else W.bowl_example(mesh=W.refined_bowl(lv))
        ops = W.host_operands(w)
        A = ops["A"]
        y = ops["B"] @ ops["b_init"] + ops["b0"]
        dA = ctx.csr(A, drop_zeros=True)
        dy = ctx.vector(y)
        for orth, name in ((lib.ORTH_MGS, "mgs"), (lib.ORTH_CGS2, "cgs2"), (lib.ORTH_CGS2_FUSED, "cgs2f")):
            for rep in range(2):
                x = ctx.vector(y.size)
                st, hist = lib.gmres_solve(dA, dy, x, pscale=ops["pscale"], atol=1e-6, rtol=1e-6, itmax=3000,
                                           orth=orth, history=4096)
            print(f"h={0.08 / 2 ** lv:g} N={A.shape[0]} {name:6s}: {1e3 * st.device_ms / st.niter:7.2f} us/iter "
                  f"({st.niter} its, rnorm/rnorm0 = {st.rnorm / st.rnorm0:.3e})", flush=True)


if __name__ == "__main__":
    main()
