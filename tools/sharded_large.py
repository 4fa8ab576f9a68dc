#!/usr/bin/env python
"""Inversion-only GMRES(20) on a refined bowl3D mesh, sharded over the ranks of a torchrun launch
(BASELINE configs[2] with --level 1 = h 0.04, configs[4]'s size with --level 2 = h 0.02).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node=N --master-addr 127.0.0.1 \
        --master-port 29551 tools/sharded_large.py --level 2 --iters 400
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=1)
    ap.add_argument("--iters", type=int, default=400)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    from bench import gmres_bytes, peaks
    from nupgcm_b200 import lib, workloads as W
    from nupgcm_b200.inversion import permuted_inversion_system
    from nupgcm_b200.sharding import torch_comm
    t0 = time.time()
    w = W.bowl_example(mesh=W.refined_bowl(args.level))
    fe = w.fe_data()
    A, B, b0, pscale = permuted_inversion_system(fe, w.params, w.forcings)
    y = B @ fe.spaces.B.interpolate(w.b0)[0][fe.dofs.p_b] + b0
    n = A.shape[0]
    arch, comm = torch_comm(n)
    ctx = arch.ctx
    dA = ctx.csr(A, drop_zeros=True)
    nnz = dA.info()["nnz_stored"]
    del A, B
    if world > 1:
        dA.shard(comm)
    dy = ctx.vector(y)
    if rank == 0:
        print(f"h = {0.08 / 2 ** args.level:g}: N = {n}, non-zeros {nnz}, {world} rank(s), set-up {time.time() - t0:.1f} s", flush=True)
    peak, _ = peaks()
    for orth, name in ((lib.ORTH_CGS2_FUSED, "cgs2f"),):
        for its in (40, args.iters):
            x = ctx.vector(n)
            dist.barrier()
            st, _ = lib.gmres_solve(dA, dy, x, pscale=pscale, atol=0, rtol=1e-30, itmax=its, orth=orth)
        t = torch.tensor([st.device_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        gbs = gmres_bytes(n, nnz, st.niter) / (ms * 1e-3) / 1e9
        if rank == 0:
            info = dA.shard_info(0) if world > 1 else {"halo_rows": 0}
            print(f"k_gmres {name} x{world}: {1e3 * ms / st.niter:8.1f} us/iter  {gbs:8.1f} GB/s algorithmic "
                  f"({gbs / (peak * world):.2f} of the measured HBM peak of {world} GPU(s)); rank-0 halo rows {info['halo_rows']}",
                  flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
