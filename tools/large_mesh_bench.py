"""Krylov / SpMV throughput on refined bowl3D meshes, where the matrix no longer fits on chip
(BASELINE configs 3 and 5: h = 0.04 and h = 0.02, obtained by refining the shipped h = 0.08 mesh).

    python tools/large_mesh_bench.py [levels ...]      (default: 1 -> h = 0.04)
"""
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
    levels = [int(v) for v in sys.argv[1:]] or [1]
    ctx = GPU(0).ctx
    peak, _ = peaks()
    for lv in levels:
        t0 = time.time()
        w = W.bowl_example(mesh=W.refined_bowl(lv))
        ops = W.host_operands(w)
        A = ops["A"]
        y = ops["B"] @ ops["b_init"] + ops["b0"]
        print(f"== h = {0.08 / 2 ** lv:g}: N = {A.shape[0]}, nnz stored {A.nnz}, non-zero "
              f"{np.count_nonzero(A.data)} (host set-up {time.time() - t0:.1f} s)", flush=True)
        dA = ctx.csr(A, drop_zeros=True)
        nnz = dA.info()["nnz_stored"]
        n = A.shape[0]
        dx, dy = ctx.vector(np.random.default_rng(0).uniform(-1, 1, n)), ctx.vector(n)
        for _ in range(3):
            dA.spmv(dx, dy)
        ctx.synchronize()
        ctx.timer_start()
        reps = 20
        for _ in range(reps):
            dA.spmv(dx, dy)
        ms = ctx.timer_stop() / reps
        gbs = spmv_bytes(n, nnz) / (ms * 1e-3) / 1e9
        print(f"stand-alone k_spmv: {ms * 1e3:8.1f} us  {gbs:7.1f} GB/s  ({gbs / peak:.2f} of measured HBM peak)", flush=True)
        dyv = ctx.vector(y)
        for tma in (("1",) if lv >= 2 else ("1", "0")):
            os.environ["NUPGCM_STREAM_TMA"] = tma
            for orth, name in ((lib.ORTH_CGS2_FUSED, "cgs2f"), (lib.ORTH_CGS2, "cgs2"), (lib.ORTH_MGS, "mgs")):
                x = ctx.vector(n)
                lib.gmres_solve(dA, dyv, x, pscale=ops["pscale"], atol=0, rtol=1e-30, itmax=40, orth=orth)
                x = ctx.vector(n)
                st, _ = lib.gmres_solve(dA, dyv, x, pscale=ops["pscale"], atol=0, rtol=1e-30, itmax=400, orth=orth)
                us = 1e3 * st.device_ms / st.niter
                gbs = gmres_bytes(n, nnz, st.niter) / (st.device_ms * 1e-3) / 1e9
                print(f"k_gmres tma={tma} {name:5s}: {us:8.1f} us/iter  {gbs:7.1f} GB/s algorithmic "
                      f"({gbs / peak:.2f} of measured HBM peak)", flush=True)
        os.environ.pop("NUPGCM_STREAM_TMA")


if __name__ == "__main__":
    main()
