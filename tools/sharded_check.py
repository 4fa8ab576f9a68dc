#!/usr/bin/env python
"""One process per GPU (launch with torchrun): sharded CG / GMRES over CUDA-IPC arenas against the
single-GPU solver of the same process, on the bowl3D operands.  Prints timings per iteration and
'sharded_check ok' on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node=N --master-addr 127.0.0.1 \
        --master-port 29531 tools/sharded_check.py [--h 0.1] [--itmax 600]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--h", type=float, default=0.1)
    ap.add_argument("--itmax", type=int, default=600)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()

    from nupgcm_b200 import lib, workloads as W
    from nupgcm_b200.sharding import torch_comm
    w = W.bowl_example(h=args.h) if args.h != 0.1 else W.bowl_mixing()
    ops = W.host_operands(w)
    A = ops["A"].tocsr()
    E = (ops["M"] + 0.05 * (ops["Kh"] + ops["Kv"])).tocsr()
    arch, comm = torch_comm(A.shape[0])
    ctx = arch.ctx
    rng = np.random.default_rng(5)
    bA = rng.uniform(-1, 1, A.shape[0])
    bE = rng.uniform(-1, 1, E.shape[0])
    dinv = 1.0 / E.diagonal()

    def rel(a, b):
        return float(np.linalg.norm(a - b) / np.linalg.norm(b))

    ok = True
    # ---- CG
    outs = {}
    for mode in ("single", "sharded"):
        dE = ctx.csr(E)
        if mode == "sharded":
            dE.shard(comm)
        x = ctx.vector(E.shape[0])
        st, hist = lib.cg_solve(dE, ctx.vector(bE), x, dinv=ctx.vector(dinv), atol=0.0, rtol=1e-12, history=4096)
        outs[mode] = (st.niter, hist, x.download(), st.device_ms)
        dist.barrier()
    s, m = outs["single"], outs["sharded"]
    ok &= abs(s[0] - m[0]) <= 1 and rel(m[2], s[2]) < 1e-9
    ok &= float(np.linalg.norm(E @ m[2] - bE) / np.linalg.norm(bE)) < 1e-10
    if rank == 0:
        print(f"CG    n={E.shape[0]} single {s[0]} its {1e3*s[3]/s[0]:.2f} us/it | sharded x{world} {m[0]} its "
              f"{1e3*m[3]/m[0]:.2f} us/it | rel diff {rel(m[2], s[2]):.2e}", flush=True)
    # ---- GMRES (both orthogonalisations)
    for orth, name in ((lib.ORTH_MGS, "mgs"), (lib.ORTH_CGS2, "cgs2"), (lib.ORTH_CGS2_FUSED, "cgs2f")):
        outs = {}
        for mode in ("single", "sharded"):
            dA = ctx.csr(A, drop_zeros=True)
            if mode == "sharded":
                dA.shard(comm)
            x = ctx.vector(A.shape[0])
            st, hist = lib.gmres_solve(dA, ctx.vector(bA), x, pscale=ops["pscale"], atol=1e-6, rtol=1e-6,
                                       itmax=args.itmax, memory=20, orth=orth, history=8192)
            outs[mode] = (st.niter, hist, x.download(), st.device_ms)
            dist.barrier()
        s, m = outs["single"], outs["sharded"]
        k = min(len(s[1]), len(m[1]), 300)
        ok &= s[0] == m[0] and bool(np.allclose(s[1][:k], m[1][:k], rtol=1e-6)) and rel(m[2], s[2]) < 1e-6
        if rank == 0:
            print(f"GMRES {name} N={A.shape[0]} single {s[0]} its {1e3*s[3]/s[0]:.2f} us/it | sharded x{world} "
                  f"{m[0]} its {1e3*m[3]/m[0]:.2f} us/it | rel diff {rel(m[2], s[2]):.2e}", flush=True)
    # every rank must hold the same solution bits
    t = torch.from_numpy(outs["sharded"][2]).cuda()
    lst = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(lst, t)
    ok &= all(bool(torch.equal(lst[0], u)) for u in lst)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("sharded_check ok" if int(flag) == 1 else "sharded_check FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
