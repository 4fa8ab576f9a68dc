"""The streaming SpMV engine of the persistent solvers alone (nupgcm_diag_stream_spmv): GB/s against
12 nnz + 20 n algorithmic bytes, checked against SciPy, with the two timing-experiment modes.

    python tools/spmv_engine_bench.py [--level 1] [--reps 50] [--fmax 4096] [--modes 0 1 2]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import spmv_bytes, peaks                  # noqa: E402
from nupgcm_b200 import workloads as W               # noqa: E402
from nupgcm_b200.architectures import GPU            # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=1)
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--fmax", type=int, nargs="*", default=[2560])
    ap.add_argument("--modes", type=int, nargs="*", default=[0, 1, 2])
    args = ap.parse_args()
    ctx = GPU(0).ctx
    peak, _ = peaks()
    t0 = time.time()
    w = W.bowl_example(mesh=W.refined_bowl(args.level)) if args.level > 0 else W.bowl_example()
    fe = w.fe_data()
    from nupgcm_b200.inversion import permuted_inversion_system
    A, _, _, _ = permuted_inversion_system(fe, w.params, w.forcings)
    n = A.shape[0]
    print(f"== level {args.level}: N = {n}, nnz stored {A.nnz} (host set-up {time.time() - t0:.1f} s)", flush=True)
    xh = np.random.default_rng(0).uniform(-1, 1, n)
    ref = A @ xh
    os.environ["NUPGCM_RESIDENT"] = "0"
    for fmax in args.fmax:
        os.environ["NUPGCM_STREAM_FMAX"] = str(fmax)
        dA = ctx.csr(A, drop_zeros=True)
        nnz = dA.info()["nnz_stored"]
        dx, dy = ctx.vector(xh), ctx.vector(n)
        dA.stream_spmv(dx, dy, reps=2)
        err = np.linalg.norm(dy.download() - ref) / np.linalg.norm(ref)
        for mode in args.modes:
            dA.stream_spmv(dx, dy, reps=3, mode=mode)
            us = dA.stream_spmv(dx, dy, reps=args.reps, mode=mode)
            gbs = spmv_bytes(n, nnz) / (us * 1e-6) / 1e9
            what = {0: "product", 1: "copy pipeline only", 2: "no footprint gather"}[mode]
            print(f"fmax {fmax} mode {mode} ({what:19s}): {us:8.1f} us  {gbs:7.1f} GB/s  ({gbs / peak:.3f} of measured HBM peak)"
                  + (f"  rel err vs SciPy {err:.1e}" if mode == 0 else ""), flush=True)
        del dA


if __name__ == "__main__":
    main()
