import os, sys, numpy as np
sys.path.insert(0, "/root/repo")
case, fmax = sys.argv[1], sys.argv[2]
os.environ["NUPGCM_RESIDENT"] = "0"; os.environ["NUPGCM_GRID"] = fmax
from nupgcm_b200 import lib, workloads as W
from nupgcm_b200.architectures import GPU
ctx = GPU(0).ctx
w = W.bowl_mixing(); ops = W.host_operands(w)
A = ops["A"]; n = A.shape[0]
rng = np.random.default_rng(0); xh = rng.uniform(-1, 1, n)
dA = ctx.csr(A, drop_zeros=True)
try:
    if case == "spmv":
        dx, dy = ctx.vector(xh), ctx.vector(n)
        for mode in (1, 0):
            us = dA.stream_spmv(dx, dy, reps=2, mode=mode)
            print(case, fmax, "mode", mode, "us", us, "err", np.linalg.norm(dy.download() - A @ xh) / np.linalg.norm(A @ xh), flush=True)
    else:
        x = ctx.vector(n)
        st, h = lib.gmres_solve(dA, ctx.vector(xh), x, pscale=ops["pscale"], atol=0, rtol=1e-30, itmax=45, orth=lib.ORTH_MGS, history=64)
        print(case, fmax, "niter", st.niter, "ms", st.device_ms, h[:3], flush=True)
except Exception as e:
    print(case, fmax, "ERROR", e, flush=True)
