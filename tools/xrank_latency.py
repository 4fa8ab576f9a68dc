#!/usr/bin/env python
"""Latency of the inter-rank primitives (run under torchrun, >= 2 ranks): flag ping-pong over
NVLink between rank 0 and every other rank, and µs per GMRES/CG iteration of the sharded solvers
for both reduction modes (NUPGCM_XMODE)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    from nupgcm_b200 import lib, workloads as W
    from nupgcm_b200.sharding import torch_comm
    w = W.bowl_example(h=0.08)
    ops = W.host_operands(w)
    A = ops["A"].tocsr()
    arch, comm = torch_comm(A.shape[0])
    ctx = arch.ctx
    for b in range(1, world):
        for variant, name in ((0, "relaxed.sys"), (1, "release/acquire.sys"), (2, "fence.sys+relaxed")):
            dist.barrier()
            us = comm.xping(0, b, variant, 5000)
            if rank == 0:
                print(f"flag over NVLink rank0<->rank{b} {name:20s}: {us:6.2f} us one way", flush=True)
    for xmode in ("1", "0"):
        os.environ["NUPGCM_XMODE"] = xmode
        for count, publish in ((1, 0), (1, 1), (1, 2), (10, 0), (20, 0)):
            dist.barrier()
            us = comm.xreduce(count, publish, 3000)
            if rank == 0:
                print(f"xmode={xmode} reduction of {count:2d} value(s), publish={int(publish)}: {us:6.2f} us ({world} ranks)", flush=True)
    y = ops["B"] @ ops["b_init"] + ops["b0"]
    dA = ctx.csr(A, drop_zeros=True).shard(comm)
    dy = ctx.vector(y)
    for xmode in ("1", "0"):
        os.environ["NUPGCM_XMODE"] = xmode
        for orth, name in ((lib.ORTH_CGS2, "cgs2"), (lib.ORTH_MGS, "mgs")):
            for rep in range(2):
                x = ctx.vector(A.shape[0])
                dist.barrier()
                st, _ = lib.gmres_solve(dA, dy, x, pscale=ops["pscale"], atol=0, rtol=1e-30, itmax=2000, orth=orth)
            if rank == 0:
                print(f"xmode={xmode} gmres {name}: {1e3 * st.device_ms / st.niter:7.2f} us/iter ({world} ranks, h=0.08)", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
