"""Streamed-matrix regime of the persistent solvers (refined bowl3D meshes: the matrix slice of a CTA
does not fit in shared memory): stand-alone SpMV, GMRES(20) per-iteration time for the three
orthogonalisations, and the in-kernel phase split (CTA 0's clock: SpMV / local vector work /
reduction wait / scalar recurrences).

    python tools/stream_bench.py [--level 1] [--iters 400] 
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import gmres_bytes, spmv_bytes, peaks    # noqa: E402
from nupgcm_b200 import lib                          # noqa: E402
from nupgcm_b200 import workloads as W               # noqa: E402
from nupgcm_b200.architectures import GPU            # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=1)
    ap.add_argument("--iters", type=int, default=400)
    ap.add_argument("--orth", nargs="*", default=["mgs", "cgs2f"])
    ap.add_argument("--no-spmv", action="store_true")
    args = ap.parse_args()
    ctx = GPU(0).ctx
    peak, _ = peaks()
    t0 = time.time()
    w = W.bowl_example(mesh=W.refined_bowl(args.level))
    fe = w.fe_data()
    from nupgcm_b200.inversion import permuted_inversion_system
    A, B, b0, pscale = permuted_inversion_system(fe, w.params, w.forcings)
    b_init = fe.spaces.B.interpolate(w.b0)[0][fe.dofs.p_b]
    y = B @ b_init + b0
    n = A.shape[0]
    print(f"== h = {0.08 / 2 ** args.level:g}: N = {n}, nnz stored {A.nnz} (host set-up {time.time() - t0:.1f} s)", flush=True)
    names = {"mgs": lib.ORTH_MGS, "cgs2": lib.ORTH_CGS2, "cgs2f": lib.ORTH_CGS2_FUSED}
    for fmax in ["-"]:
        t0 = time.time()
        dA = ctx.csr(A, drop_zeros=True)
        nnz = dA.info()["nnz_stored"]
        if not args.no_spmv:
            dx, dy = ctx.vector(np.random.default_rng(0).uniform(-1, 1, n)), ctx.vector(n)
            for _ in range(3):
                dA.spmv(dx, dy)
            ctx.synchronize()
            ctx.timer_start()
            for _ in range(20):
                dA.spmv(dx, dy)
            ms = ctx.timer_stop() / 20
            gbs = spmv_bytes(n, nnz) / (ms * 1e-3) / 1e9
            print(f"stand-alone k_spmv: {ms * 1e3:8.1f} us  {gbs:7.1f} GB/s  ({gbs / peak:.2f} of measured HBM peak)", flush=True)
        dyv = ctx.vector(y)
        x = ctx.vector(n)
        lib.gmres_solve(dA, dyv, x, pscale=pscale, atol=0, rtol=1e-30, itmax=20, orth=lib.ORTH_MGS)
        print(f"nnz {nnz}, device tables + first solve {time.time() - t0:.1f} s", flush=True)
        for name in args.orth:
            for prof in ("0", "1"):
                os.environ["NUPGCM_PROFILE"] = prof
                x = ctx.vector(n)
                st, _ = lib.gmres_solve(dA, dyv, x, pscale=pscale, atol=0, rtol=1e-30, itmax=args.iters, orth=names[name])
                us = 1e3 * st.device_ms / st.niter
                gbs = gmres_bytes(n, nnz, st.niter) / (st.device_ms * 1e-3) / 1e9
                if prof == "0":
                    print(f"k_gmres {name:5s}: {us:8.1f} us/iter  {gbs:7.1f} GB/s algorithmic "
                          f"({gbs / peak:.3f} of measured HBM peak)  rnorm/rnorm0 {st.rnorm / st.rnorm0:.3e}", flush=True)
                else:
                    pf = [us * f for f in st.phase_frac]
                    print(f"        phases (profiled build, {us:.1f} us/iter): spmv {pf[0]:.1f}  local {pf[1]:.1f}  "
                          f"reduce {pf[2]:.1f}  scalar {pf[3]:.1f} us   SM {st.sm_mhz:.0f} MHz   "
                          f"spmv alone = {spmv_bytes(n, nnz) / (pf[0] * 1e-6) / 1e9:.0f} GB/s", flush=True)
            os.environ.pop("NUPGCM_PROFILE")
        del dA


if __name__ == "__main__":
    main()
