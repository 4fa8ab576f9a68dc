"""Worker of tests/test_sharded_host.py (world size 2, gloo, CPU only): executes the sharded-solve
LAYOUT computed by the library's host-only planner (nupgcm_shard_plan: internal ordering, row
blocks, halo push ranges) with NumPy ranks that exchange exactly the planned halo rows
(torch.distributed send/recv) and all-reduce their dot products — the communication pattern the
CUDA kernels implement with peer-memory stores.  A Jacobi-CG run on that layout must reproduce the
serial oracle's residual history."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    from conftest import workload
    from nupgcm_b200 import lib
    from oracle import krylov
    _, ops = workload("bowl_mixing", dim=2)
    A = (ops["M"] + 0.05 * (ops["Kh"] + ops["Kv"])).tocsr()
    n = A.shape[0]
    plan = lib.shard_plan(A, world, grid_per_rank=16)
    perm, rb = plan["perm"], plan["row_begin"]
    P = A[perm][:, perm].tocsr()
    r0, r1 = int(rb[rank]), int(rb[rank + 1])
    mine = P[r0:r1]                                        # this rank's row block
    dinv = 1.0 / P.diagonal()[r0:r1]
    rng = np.random.default_rng(0)
    b_full = rng.uniform(-1, 1, n)
    b = b_full[perm][r0:r1]

    def exchange(v_own):
        """Full-length vector holding own rows + the planned halo rows (NaN elsewhere)."""
        full = np.full(n, np.nan)
        full[r0:r1] = v_own
        reqs = []
        for p in range(world):
            if p == rank:
                continue
            lo, hi = int(plan["halo_lo"][p, rank]), int(plan["halo_hi"][p, rank])   # my rows -> p
            if hi > lo:
                reqs.append(dist.isend(torch.from_numpy(full[lo:hi].copy()), p))
        for p in range(world):
            if p == rank:
                continue
            lo, hi = int(plan["halo_lo"][rank, p]), int(plan["halo_hi"][rank, p])   # p's rows -> me
            if hi > lo:
                t = torch.empty(hi - lo, dtype=torch.float64)
                dist.recv(t, p)
                full[lo:hi] = t.numpy()
        for q in reqs:
            q.wait()
        return full

    def spmv(v_own):
        y = mine @ np.nan_to_num(exchange(v_own), nan=np.inf)   # an unplanned column would poison y
        assert np.all(np.isfinite(y)), "halo plan does not cover the rank's column footprint"
        return y

    def dot(a, c):
        t = torch.tensor([float(a @ c)], dtype=torch.float64)
        dist.all_reduce(t)
        return float(t[0])

    # Jacobi-CG (Krylov.jl order) on the sharded layout
    x = np.zeros(r1 - r0)
    r = b - spmv(x)
    z = dinv * r
    p = z.copy()
    gamma = dot(r, z)
    hist = [np.sqrt(gamma)]
    for _ in range(25):
        Ap = spmv(p)
        alpha = gamma / dot(p, Ap)
        x += alpha * p
        r -= alpha * Ap
        z = dinv * r
        g2 = dot(r, z)
        hist.append(np.sqrt(g2))
        p = z + (g2 / gamma) * p
        gamma = g2
    _, so = krylov.cg(A, b_full, x0=np.zeros(n), M=1.0 / A.diagonal(), atol=0.0, rtol=0.0, itmax=25)
    ref = np.asarray(so.residuals[:26])
    assert np.allclose(hist, ref, rtol=1e-9), (hist[:4], ref[:4])
    # ownership covers every row exactly once
    cnt = torch.tensor([r1 - r0], dtype=torch.int64)
    dist.all_reduce(cnt)
    assert int(cnt[0]) == n
    if rank == 0:
        print("gloo sharded layout ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
