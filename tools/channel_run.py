"""BASELINE configs[3]: channel_basin with wind and surface-flux forcing on the GPUs of one box.

    torchrun --nproc-per-node 8 tools/channel_run.py --cells 32 64 8 --steps 30

The reference's production set-up (scratch/run.jl:111-163): channel_basin_flat mesh (x-periodic channel for
y <= -1/2, walled basin north of it; here the structured substitute box, meshes/channel_basin_flat.jl needs
gmsh), P2-P1 flow with FIRST-order buoyancy (b_order = 1), wind stress + surface buoyancy flux,
BDF1(adaptive = true) with the CFL step computed on the device every step, the convection parameterisation
(Kv and its right-hand sides rebuilt on the device every step) and the eddy parameterisation (friction block
of the inversion matrix rebuilt every 10 steps).  Both Krylov solves are row-block sharded over the ranks;
every rank assembles the (small) host operands itself.  Reference loop: src/model.jl:128-209."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", dest="n", type=int, nargs=3, default=[32, 64, 8], help="box cells in x, y, z")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--orth", default="mgs", choices=["mgs", "cgs2f"])
    ap.add_argument("--b-order", type=int, default=1)
    ap.add_argument("--cfl", type=float, default=0.8, help="scratch/run.jl:163 uses 0.8 on its h = 1e-2 mesh")
    ap.add_argument("--check", action="store_true", help="rank 0 repeats the run on its GPU alone and compares")
    ap.add_argument("--out", default="gpurun_out/channel_run.json")
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

    t0 = time.perf_counter()
    w = W.with_b_order(W.channel_basin_box(n=tuple(args.n), periodic=True), args.b_order)
    w.timestepper_kwargs = dict(w.timestepper_kwargs, CFL_factor=args.cfl)
    ops = W.host_operands(w)
    t_host = time.perf_counter() - t0
    fe = w.fe_data()
    if world > 1:
        from nupgcm_b200.sharding import torch_comm
        arch, comm = torch_comm(max_n=ops["A"].shape[0] + 16)
    else:
        arch = npg.GPU(local)
    orth = {"mgs": lib.ORTH_MGS, "cgs2f": lib.ORTH_CGS2_FUSED}[args.orth]

    def make(a):
        inv = npg.InversionToolkit(a, ops["A"], ops["pscale"], ops["B"], ops["b0"], orth=orth, drop_zeros=False, history=False)
        ts = w.timestepper()
        evo = npg.EvolutionToolkit(a, ops, w.params, w.forcings, ts, history=False)
        m = npg.Model(a, w.params, w.forcings, fe, inv, evo, ts, tables=ops["tables"])
        m.xb.upload(ops["b_init"])
        return m, ts

    def barrier():
        arch.ctx.synchronize()
        if dist is not None:
            dist.barrier()

    t1 = time.perf_counter()
    m, ts = make(arch)
    barrier()
    t_dev = time.perf_counter() - t1
    dts = []
    t2 = time.perf_counter()
    for _ in range(args.steps):
        npg.run_(m, n_steps=1, resume=True)
        dts.append(ts.Δt)
    barrier()
    t_run = time.perf_counter() - t2
    log = m.step_log
    g_it = np.array([r["gmres_iters"] for r in log], dtype=float)
    g_ms = np.array([r["gmres_ms"] for r in log], dtype=float)
    c_it = np.array([r["cg_iters"] for r in log], dtype=float)
    c_ms = np.array([r["cg_ms"] for r in log], dtype=float)
    info = m.inversion.solver.A.info()
    free, total = arch.ctx.mem_status()
    out = {"workload": f"channel_basin box {args.n[0]}x{args.n[1]}x{args.n[2]} cells x 6 tetrahedra, channel y<=-1/2 periodic in x, "
                       f"b_order={args.b_order}, wind + surface flux, BDF1(adaptive, CFL {args.cfl}), convection rebuild every "
                       "step, eddy rebuild every 10 steps (BASELINE configs[3], scratch/run.jl:111-163)",
           "n_gpus": world, "orth": args.orth, "N": info["n_rows"], "nnz": info["nnz_stored"], "nb": int(ops["nb"]),
           "steps": args.steps, "timesteps_per_s": args.steps / t_run, "s_per_step": t_run / args.steps,
           "first_step_gmres_iters": float(g_it[0]), "gmres_per_step_mean_after_first": float(g_it[1:].mean()) if len(g_it) > 1 else None,
           "gmres_us_per_iter": float(1e3 * g_ms.sum() / max(g_it.sum(), 1)),
           "cg_per_step_mean": float(c_it.mean()), "cg_us_per_iter": float(1e3 * c_ms.sum() / max(c_it.sum(), 1)),
           "solve_share_of_run": float((g_ms.sum() + c_ms.sum()) * 1e-3 / t_run),
           "dt_first_last": [dts[0], dts[-1]], "t_end": ts.t,
           "all_solved": bool(all(r["gmres_solved"] and r["cg_solved"] for r in log)),
           "per_step": {"gmres_iters": [int(v) for v in g_it], "gmres_solved": [bool(r["gmres_solved"]) for r in log],
                        "cg_iters": [int(v) for v in c_it], "cg_solved": [bool(r["cg_solved"]) for r in log],
                        "dt": [float(v) for v in dts]},
           "u_max_last": log[-1]["u_max"], "b_max_last": log[-1]["b_max"],
           "setup_s": {"host_assembly_per_rank": t_host, "device_tables_and_upload": t_dev},
           "device_memory_used_gb_rank0": (total - free) / 2 ** 30}
    if args.check and rank == 0 and world > 1:
        single, ts1 = make(npg.GPU(local))
        for _ in range(args.steps):
            npg.run_(single, n_steps=1, resume=True)
        rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))      # noqa: E731
        out["sharded_vs_single"] = {"rel_diff_u_p": rel(m.inversion.solver.x.download(), single.inversion.solver.x.download()),
                                    "rel_diff_b": rel(m.xb.download(), single.xb.download()),
                                    "gmres_iters_single": [float(r["gmres_iters"]) for r in single.step_log][:5],
                                    "gmres_iters_sharded": [float(v) for v in g_it[:5]]}
    barrier()
    if rank == 0:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        json.dump(out, open(args.out, "w"), indent=1)
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
