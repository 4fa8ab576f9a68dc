"""The streaming SpMV engine of the persistent solvers alone (nupgcm_diag_stream_spmv): GB/s against
12 nnz + 20 n algorithmic bytes, checked against SciPy, with the two timing-experiment modes.

    python tools/spmv_engine_bench.py [--level 1] [--reps 50] [--modes 0 1 2]
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
    for fmax in ["-"]:
        dA = ctx.csr(A, drop_zeros=True)
        nnz = dA.info()["nnz_stored"]
        dx, dy = ctx.vector(xh), ctx.vector(n)
        dA.stream_spmv(dx, dy, reps=2)
        err = np.linalg.norm(dy.download() - ref) / np.linalg.norm(ref)
        if 3 in args.modes:
            us, cyc = dA.stream_spmv(dx, dy, reps=args.reps, mode=3)
            names = ["wait footprint", "wait ring", "tables+misc", "full positions", "jagged ends", "callback", "closing barrier"]
            tot = cyc[:, :, :8].sum(axis=2)
            print(f"mode 3: {us:.1f} us per product; per-warp SM cycles per product (mean over all warps | slowest warp of CTA 0):")
            w = np.unravel_index(np.argmax(cyc[0, :, :6].sum(axis=1)), (11,))[0]
            for k, nm in enumerate(names):
                print(f"    {nm:16s} {cyc[:, :, k].mean():9.0f} | {cyc[0, w, k]:9.0f}")
            print(f"    {'ring refills':16s} {cyc[:, :, 7].mean():9.0f} | {cyc[0, w, 7]:9.0f}")
            print(f"    {'total':16s} {tot.mean():9.0f} | {tot[0, w]:9.0f}   busy (no closing barrier): min {cyc[:, :, :6].sum(axis=2).min():.0f} max {cyc[:, :, :6].sum(axis=2).max():.0f}", flush=True)
        for mode in [m for m in args.modes if m != 3]:
            dA.stream_spmv(dx, dy, reps=3, mode=mode)
            us = dA.stream_spmv(dx, dy, reps=args.reps, mode=mode)
            gbs = spmv_bytes(n, nnz) / (us * 1e-6) / 1e9
            what = {0: "product", 1: "copy pipeline only", 2: "no footprint gather", 4: "compute only, no copy"}[mode]
            print(f"mode {mode} ({what:19s}): {us:8.1f} us  {gbs:7.1f} GB/s  ({gbs / peak:.3f} of measured HBM peak)"
                  + (f"  rel err vs SciPy {err:.1e}" if mode == 0 else ""), flush=True)
        del dA


if __name__ == "__main__":
    main()
