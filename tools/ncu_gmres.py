"""One bounded GMRES(20) solve on the h = 0.04 inversion system for ncu (application replay: the persistent
cooperative kernel must not be kernel-replayed).  python tools/ncu_gmres.py [mgs|cgs2f] [iters]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nupgcm_b200 import lib, workloads as W          # noqa: E402
from nupgcm_b200.architectures import GPU            # noqa: E402
from nupgcm_b200.inversion import permuted_inversion_system   # noqa: E402

orth = {"mgs": lib.ORTH_MGS, "cgs2f": lib.ORTH_CGS2_FUSED}[sys.argv[1] if len(sys.argv) > 1 else "mgs"]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
ctx = GPU(0).ctx
w = W.bowl_example(mesh=W.refined_bowl(1))
fe = w.fe_data()
A, B, b0, pscale = permuted_inversion_system(fe, w.params, w.forcings)
y = B @ fe.spaces.B.interpolate(w.b0)[0][fe.dofs.p_b] + b0
dA = ctx.csr(A, drop_zeros=True)
x = ctx.vector(y.size)
lib.gmres_solve(dA, ctx.vector(y), ctx.vector(y.size), pscale=pscale, atol=0, rtol=1e-30, itmax=20, orth=orth)   # launch 0: warm-up
st, _ = lib.gmres_solve(dA, ctx.vector(y), x, pscale=pscale, atol=0, rtol=1e-30, itmax=iters, orth=orth)
print(f"{st.niter} iterations, {1e3 * st.device_ms / st.niter:.1f} us/iter")
