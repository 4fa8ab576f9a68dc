"""Per-iteration cost of the persistent Krylov kernels under different launch variants.

    python tools/krylov_microbench.py [h ...]

Runs GMRES(20) for a fixed 2000 iterations and CG for 200 on the bowl3D operators and prints
µs/iteration and the algorithmic GB/s for: cooperative grid size, SM-resident vs streaming matrix,
MGS vs CGS2, stored vs dropped explicit zeros.  Variants are switched through the NUPGCM_*
environment knobs the library reads at each solve.  Also prints the latency of the grid-wide
reduction primitive."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import cg_bytes, gmres_bytes          # noqa: E402
from nupgcm_b200 import lib                       # noqa: E402
from nupgcm_b200 import workloads as W            # noqa: E402
from nupgcm_b200.architectures import GPU         # noqa: E402

GRIDS = (148, 112, 96, 74, 64, 48, 32, 24, 16)


def setenv(**kw):
    for k, v in kw.items():
        if v is None or v == "":
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)


def main():
    hs = [float(v) for v in sys.argv[1:]] or [0.1, 0.08]
    ctx = GPU(0).ctx
    for grid in (148, 96, 64, 32, 16, 2):
        for depth in (1, 2, 4):
            setenv(NUPGCM_POLL_DEPTH=depth)
            print(f"reduce latency grid={grid:3d} depth={depth}: " + "  ".join(
                f"mode{m}={ctx.reduce_latency(m, 5000, grid, 512):6.2f}us" for m in (0, 1, 2)), flush=True)
    setenv(NUPGCM_POLL_DEPTH=None)
    for h in hs:
        w = W.bowl_example(h=h)
        ops = W.host_operands(w)
        A = ops["A"]
        y = ops["B"] @ ops["b_init"] + ops["b0"]
        Ae = (ops["M"] + 1e-4 * (ops["Kh"] + ops["Kv"])).tocsr()
        ye = np.random.default_rng(0).uniform(-1, 1, Ae.shape[0])
        print(f"== h={h}: N={A.shape[0]} nnz={A.nnz} (nonzero {np.count_nonzero(A.data)}), "
              f"nb={Ae.shape[0]} nnz_evol={Ae.nnz}", flush=True)
        for drop in (True, False):
            dA = ctx.csr(A, drop_zeros=drop)
            nnz = dA.info()["nnz_stored"]
            dy = ctx.vector(y)
            for grid in GRIDS:
                for res in ("1", "0"):
                    out = []
                    for orth, oname in ((lib.ORTH_MGS, "mgs"), (lib.ORTH_CGS2, "cgs2")):
                        setenv(NUPGCM_RESIDENT=res, NUPGCM_GRID=grid)
                        x = ctx.vector(y.size)
                        lib.gmres_solve(dA, dy, x, pscale=ops["pscale"], atol=0, rtol=1e-30, itmax=100, orth=orth)
                        x = ctx.vector(y.size)
                        st, _ = lib.gmres_solve(dA, dy, x, pscale=ops["pscale"], atol=0, rtol=1e-30,
                                                itmax=1000, orth=orth)
                        us = 1e3 * st.device_ms / st.niter
                        gbs = gmres_bytes(A.shape[0], nnz, st.niter) / (st.device_ms * 1e-3) / 1e9
                        out.append(f"{oname} {us:7.2f} us/iter {gbs:7.1f} GB/s")
                    print(f"gmres drop={int(drop)} grid={grid:3d} resident={res}: " + " | ".join(out), flush=True)
        dAe = ctx.csr(Ae)
        dinv = ctx.vector(1.0 / Ae.diagonal())
        for grid in GRIDS:
            for res in ("1", "0"):
                setenv(NUPGCM_RESIDENT=res, NUPGCM_GRID=grid)
                x = ctx.vector(ye.size)
                lib.cg_solve(dAe, ctx.vector(ye), x, dinv=dinv, atol=0, rtol=1e-300, itmax=50)
                x = ctx.vector(ye.size)
                st, _ = lib.cg_solve(dAe, ctx.vector(ye), x, dinv=dinv, atol=0, rtol=1e-300, itmax=200)
                us = 1e3 * st.device_ms / max(st.niter, 1)
                gbs = cg_bytes(Ae.shape[0], Ae.nnz, st.niter) / (st.device_ms * 1e-3) / 1e9
                print(f"cg    grid={grid:3d} resident={res}: {us:7.2f} us/iter ({st.niter} its) {gbs:7.1f} GB/s",
                      flush=True)
    setenv(NUPGCM_RESIDENT=None, NUPGCM_GRID=None)


if __name__ == "__main__":
    main()
