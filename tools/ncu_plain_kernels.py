"""The plain (non-persistent) kernels of the timestep on the h = 0.04 bowl, two launches each, for one
`ncu --set full` capture (kernel replay is safe here: none of them waits on another CTA):
k_spmv (stand-alone CSR SpMV of the inversion matrix), k_elem + k_gather_elem (advection right-hand side),
k_cfl (adaptive timestep), k_kv_elem + k_gather_mat + k_gather_elem (convection rebuild), k_nu_elem +
k_gather_mat_add (eddy rebuild).  Prints CUDA-event times and algorithmic GB/s.

    python tools/ncu_plain_kernels.py [level]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import peaks, spmv_bytes                    # noqa: E402
from nupgcm_b200 import lib, workloads as W            # noqa: E402
from nupgcm_b200._forms import build_A_inversion       # noqa: E402
from nupgcm_b200.architectures import GPU              # noqa: E402

level = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ctx = GPU(0).ctx
peak, _ = peaks()
w = W.bowl_example(mesh=W.refined_bowl(level))
ops = W.host_operands(w)
fe = w.fe_data()
tb = ops["tables"]
nc, nq = tb["cell_b"].shape[0], tb["w"].size
nb, nu, N = ops["nb"], ops["nu"], ops["A"].shape[0]
rng = np.random.default_rng(0)


def timed(name, fn, nbytes, reps=int(os.environ.get("NCU_REPS", "5"))):
    fn()
    ctx.synchronize()
    ctx.timer_start()
    for _ in range(reps):
        fn()
    ms = ctx.timer_stop() / reps
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(f"{name:34s} {ms * 1e3:9.1f} us  {gbs:8.1f} GB/s algorithmic  ({gbs / peak:.2f} of measured HBM peak)", flush=True)


print(f"h = {0.08 / 2 ** level:g}: {nc} cells, nq = {nq}, nb = {nb}, nu = {nu}, N = {N}")
dA = ctx.csr(ops["A"], drop_zeros=False)
nnz = dA.info()["nnz_stored"]
x, y = ctx.vector(rng.uniform(-1, 1, N)), ctx.vector(N)
timed("k_spmv (A, explicit zeros kept)", lambda: dA.spmv(x, y), spmv_bytes(N, nnz))
mesh = lib.ElementMesh(ctx, tb)
b, bp = ctx.vector(rng.uniform(-1, 1, nb)), ctx.vector(rng.uniform(-1, 1, nb))
u, up = ctx.vector(rng.uniform(-1, 1, N)), ctx.vector(rng.uniform(-1, 1, N))
out = ctx.vector(nb)
nlb, nlu, nv = tb["cell_b"].shape[1], tb["cell_u"].shape[1], tb["bary"].shape[1]
# per cell: index tables (4 B each), gradients + volume, gathered fields (b, b_prev, u, u_prev: 8 B each), elemental
# vector written and read once more by the gather
elem_bytes = nc * (4 * (nlb + 3 * nlu) + 8 * (3 * nv + 1) + 8 * (2 * nlb + 6 * nlu) + 16 * nlb) + 8 * nb
timed("k_elem + k_gather_elem (BDF2)", lambda: mesh.rhs_adv(2, 1e-3, 2.0, b, bp, u, up, out), elem_bytes)
timed("k_cfl", lambda: mesh.cfl_dt(u, 0.8, 0.01), nc * (4 * 3 * nlu + 8 * 3 * nlu + 8))
kv_q = fe.mesh.dΩ.coefficient(w.forcings.κᵥ, slice(None))
Kv = ctx.csr(ops["Kv"])
mesh.enable_kv_rebuild(Kv, kv_q)
rv, rd = ctx.vector(nb), ctx.vector(nb)
kv_bytes = nc * (4 * nlb + 8 * (nv + 1) + 8 * nlb + 8 * nq + 16 * (nlb * nlb + 2 * nlb)) + 12 * ops["Kv"].nnz
timed("k_kv_elem + gathers (Kv, rhs)", lambda: mesh.rebuild_kv(w.params.α, w.params.N2, 1.0, 1e-3, b, Kv, rv, rd), kv_bytes)
p = fe.dofs.p_inversion
A0 = build_A_inversion(fe, w.params, 0.0)[p][:, p].tocsr()
A0.sort_indices()
f_q = fe.mesh.dΩ.coefficient(w.params.f, slice(None))
mesh.enable_nu_rebuild(dA, A0.data, f_q)
nd = 3 * nlu
nu_bytes = nc * (4 * nlb + 8 * (3 * nv + 1) + 8 * nlb + 8 * nq + 16 * nd * nd) + 20 * ops["A"].nnz
timed("k_nu_elem + k_gather_mat_add", lambda: mesh.rebuild_friction(w.params.α ** 2 * w.params.ε ** 2, w.params.α, w.params.N2,
                                                                     0.03, 10.0, 1.0, b, dA), nu_bytes)
