"""BASELINE configs[4]: long buoyancy-evolution run on the twice-refined bowl3D mesh (h = 0.02, N = 2.15 M
velocity + pressure DOFs, 127 M non-zeros), both Krylov solves row-block sharded over the GPUs of one box.

    torchrun --nproc-per-node 8 tools/long_run.py --level 2 --steps 100 [--orth cgs2f]

Reference loop: src/model.jl:128-209 (run!), set-up of examples/bowl_mixing.jl (BDF2, Δt = 1e-3, initial
inversion).  Host set-up is done ONCE: rank 0 assembles the operands (gridap_lite, ~1 minute at this size)
and leaves them in a cache directory (default /dev/shm); the other ranks load them from there.  Every rank
then hands the full matrices to the library, which keeps only its own row block in the solver tables
(`nupgcm_csr_shard`); the state vectors are replicated, the element kernels run on every rank."""
import argparse
import json
import os
import sys
import time
from types import SimpleNamespace

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def save_ops(path, ops, dofs):
    os.makedirs(path, exist_ok=True)
    for k in ("A", "B", "M", "Kh", "Kv"):
        m = ops[k].tocsr()
        np.save(os.path.join(path, k + "_data.npy"), m.data)
        np.save(os.path.join(path, k + "_indices.npy"), m.indices)
        np.save(os.path.join(path, k + "_indptr.npy"), m.indptr)
        np.save(os.path.join(path, k + "_shape.npy"), np.array(m.shape))
    vec = {k: ops[k] for k in ("b0", "rhs_diff", "rhs_flux", "rhs_m", "rhs_h", "rhs_v", "b_init")}
    np.savez(os.path.join(path, "vectors.npz"), pscale=ops["pscale"], nu=ops["nu"], np_=ops["np"], nb=ops["nb"], **vec)
    np.savez(os.path.join(path, "tables.npz"), **{k: np.asarray(v) for k, v in ops["tables"].items()})
    np.savez(os.path.join(path, "dofs.npz"), p_b=dofs.p_b, inv_p_b=dofs.inv_p_b, p_inversion=dofs.p_inversion,
             inv_p_inversion=dofs.inv_p_inversion)
    open(os.path.join(path, "done"), "w").write("ok")


def load_ops(path):
    ops = {}
    for k in ("A", "B", "M", "Kh", "Kv"):
        shape = tuple(np.load(os.path.join(path, k + "_shape.npy")))
        ops[k] = sp.csr_matrix((np.load(os.path.join(path, k + "_data.npy")), np.load(os.path.join(path, k + "_indices.npy")),
                                np.load(os.path.join(path, k + "_indptr.npy"))), shape=shape)
    v = np.load(os.path.join(path, "vectors.npz"))
    for k in ("b0", "rhs_diff", "rhs_flux", "rhs_m", "rhs_h", "rhs_v", "b_init"):
        ops[k] = v[k]
    ops["pscale"], ops["nu"], ops["np"], ops["nb"] = float(v["pscale"]), int(v["nu"]), int(v["np_"]), int(v["nb"])
    t = np.load(os.path.join(path, "tables.npz"), allow_pickle=True)
    ops["tables"] = {k: (t[k].item() if t[k].shape == () else t[k]) for k in t.files}
    d = np.load(os.path.join(path, "dofs.npz"))
    dofs = SimpleNamespace(nu=ops["nu"], np=ops["np"], nb=ops["nb"], p_b=d["p_b"], inv_p_b=d["inv_p_b"],
                           p_inversion=d["p_inversion"], inv_p_inversion=d["inv_p_inversion"])
    return ops, dofs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=2)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--orth", default="cgs2f", choices=["mgs", "cgs2f"])
    ap.add_argument("--cache", default="/dev/shm/nupgcm_ops")
    ap.add_argument("--out", default="gpurun_out/long_run.json")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        from datetime import timedelta
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=timedelta(minutes=30))

    import nupgcm_b200 as npg
    from nupgcm_b200 import lib
    from nupgcm_b200 import workloads as W

    cache = f"{args.cache}_L{args.level}"
    w = W.bowl_example(mesh=None if args.level == 0 else W.refined_bowl(args.level)) if rank == 0 else W.bowl_example()
    t0 = time.perf_counter()
    if rank == 0 and not os.path.exists(os.path.join(cache, "done")):
        ops = W.host_operands(w)
        t_build = time.perf_counter() - t0
        save_ops(cache, ops, w.fe_data().dofs)
    else:
        t_build = 0.0
    if dist is not None:
        dist.barrier()
    t1 = time.perf_counter()
    ops, dofs = load_ops(cache)
    t_load = time.perf_counter() - t1
    fe_stub = SimpleNamespace(dofs=dofs)

    if world > 1:
        from nupgcm_b200.sharding import torch_comm
        arch, comm = torch_comm(max_n=ops["A"].shape[0] + 16)
    else:
        arch = npg.GPU(local)
    ctx = arch.ctx
    t2 = time.perf_counter()
    orth = {"mgs": lib.ORTH_MGS, "cgs2f": lib.ORTH_CGS2_FUSED}[args.orth]
    inv = npg.InversionToolkit(arch, ops["A"], ops["pscale"], ops["B"], ops["b0"], orth=orth, drop_zeros=True, history=False)
    ts = w.timestepper()
    ts.t_stop = float("inf")
    evo = npg.EvolutionToolkit(arch, ops, w.params, w.forcings, ts, history=False)
    m = npg.Model(arch, w.params, w.forcings, fe_stub, inv, evo, ts, tables=ops["tables"])
    m.xb.upload(ops["b_init"])
    ctx.synchronize()
    t_dev = time.perf_counter() - t2
    free, total = ctx.mem_status()
    if dist is not None:
        dist.barrier()
    t3 = time.perf_counter()
    npg.invert_(m)                               # examples/bowl_mixing.jl:194
    ctx.synchronize()
    t_init = time.perf_counter() - t3
    init_iters = m.inversion.solver.stats.niter
    if dist is not None:
        dist.barrier()
    t4 = time.perf_counter()
    npg.run_(m, n_steps=args.steps)
    ctx.synchronize()
    if dist is not None:
        dist.barrier()
    t_run = time.perf_counter() - t4
    log = m.step_log
    g_it = np.array([r["gmres_iters"] for r in log], dtype=float)
    g_ms = np.array([r["gmres_ms"] for r in log], dtype=float)
    c_it = np.array([r["cg_iters"] for r in log], dtype=float)
    c_ms = np.array([r["cg_ms"] for r in log], dtype=float)
    info = m.inversion.solver.A.info()
    n, nnz = info["n_rows"], info["nnz_stored"]
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import gmres_bytes, peaks
    peak, _ = peaks()
    gbs = float(np.sum([gmres_bytes(n, nnz, k) for k in g_it]) / (g_ms.sum() * 1e-3) / 1e9)
    out = {"workload": f"bowl3D h={0.08 / 2 ** args.level:g} (shipped h=0.08 mesh refined {args.level}x), examples/bowl_mixing.jl "
                       f"set-up, {args.steps} BDF2 steps after the initial inversion (BASELINE configs[4])",
           "n_gpus": world, "orth": args.orth, "N": n, "nnz": nnz, "nb": dofs.nb,
           "timesteps_per_s": args.steps / t_run, "s_per_step": t_run / args.steps,
           "gmres_per_step_mean": float(g_it.mean()), "gmres_per_step_first_last": [float(g_it[0]), float(g_it[-1])],
           "gmres_us_per_iter": float(1e3 * g_ms.sum() / g_it.sum()), "cg_per_step_mean": float(c_it.mean()),
           "cg_us_per_iter": float(1e3 * c_ms.sum() / max(c_it.sum(), 1)),
           "solve_share_of_step": float((g_ms.sum() + c_ms.sum()) * 1e-3 / t_run),
           "all_solved": bool(all(r["gmres_solved"] and r["cg_solved"] for r in log)),
           "u_max_last": log[-1]["u_max"], "b_max_last": log[-1]["b_max"],
           "initial_inversion": {"gmres_iters": int(init_iters), "seconds": t_init},
           "gmres_algorithmic_gbs_all_gpus": gbs, "frac_of_hbm_peak_all_gpus": gbs / (peak * world),
           "setup_s": {"rank0_host_assembly": t_build, "load_cached_operands": t_load, "device_tables_and_upload": t_dev},
           "device_memory_used_gb_rank0": (total - free) / 2 ** 30}
    if rank == 0:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        json.dump(out, open(args.out, "w"), indent=1)
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
