"""Row-block sharded solves (multi-GPU path) on ONE GPU: the ranks live in this process and share
cuda:0's SMs (74 CTAs each for two ranks), so the whole inter-rank protocol — halo rows pushed into
the peer's arena, flagged-word reductions across ranks, the closing all-gather — runs exactly as
it does over NVLink, only with local "peer" pointers.  Checked against the single-rank solver
(same kernels, nranks = 1), the CPU oracle and SciPy direct solves.  The one-process-per-GPU
variant over CUDA IPC is exercised by test_two_processes_over_ipc when the box has >= 2 GPUs."""
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse.linalg as spla

from conftest import ROOT, workload
from nupgcm_b200 import lib
from nupgcm_b200.sharding import local_ranks, run_collective
from oracle import krylov

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _evol(ops, θ=0.05):
    return (ops["M"] + θ * (ops["Kh"] + ops["Kv"])).tocsr()


def _sharded(nranks, A, solve, drop_zeros=False, grid=None):
    """Run `solve(ctx, dA)` collectively on `nranks` ranks sharing cuda:0; returns per-rank results.

    The ranks share ONE device here, so a device-wide synchronisation on one rank's thread (cudaFree
    of a temporary vector) would wait for another rank's persistent kernel, which in turn may be
    waiting for this rank's next launch.  `solve` therefore runs under `_NoFree`: vectors created
    inside it stay alive until every rank has finished (with one process per GPU, the production
    set-up, a rank's cudaFree only ever waits for its own device)."""
    import threading
    comms = local_ranks(nranks, A.shape[0])
    if grid is not None:               # fewer CTAs per rank: set BEFORE sharding, so that the solver tables are built
        for c in comms:                # here and not inside the collective solve (cudaFree = device-wide sync)
            c.ctx.set_grid(grid)
    mats = [c.ctx.csr(A, drop_zeros=drop_zeros).shard(c) for c in comms]
    done = threading.Barrier(nranks)

    def rank_main(c, m):
        keep = _NoFree(c.ctx)
        try:
            return solve(keep, m)
        finally:
            try:
                done.wait(timeout=120)     # nobody frees anything before all ranks' kernels have ended
            except threading.BrokenBarrierError:
                pass
            keep.release()

    return run_collective([lambda c=c, m=m: rank_main(c, m) for c, m in zip(comms, mats)]), mats, comms


class _NoFree:
    """Context proxy whose vectors are kept alive until release()."""

    def __init__(self, ctx):
        self._ctx, self._keep = ctx, []

    def vector(self, data_or_n):
        v = self._ctx.vector(data_or_n)
        self._keep.append(v)
        return v

    def __getattr__(self, name):
        return getattr(self._ctx, name)

    def release(self):
        self._keep.clear()


@pytest.mark.parametrize("nranks", [2, 4])
def test_cg_sharded_matches_single_rank(ctx, nranks):
    _, ops = workload("bowl_mixing")
    A = _evol(ops)
    rng = np.random.default_rng(0)
    b = rng.uniform(-1, 1, A.shape[0])
    x0 = rng.uniform(-1, 1, A.shape[0])
    dinv = 1.0 / A.diagonal()

    def solve(c, dA):
        x = c.vector(x0)
        st, hist = lib.cg_solve(dA, c.vector(b), x, dinv=c.vector(dinv), atol=1e-6, rtol=1e-6, history=4096)
        return st.niter, bool(st.solved), hist, x.download()

    ref = solve(ctx, ctx.csr(A))
    res, mats, comms = _sharded(nranks, A, solve)
    xo, so = krylov.cg(A, b, x0=x0, M=dinv, atol=1e-6, rtol=1e-6)
    for r, (niter, solved, hist, x) in enumerate(res):
        assert solved
        assert abs(niter - ref[0]) <= 1 and abs(niter - so.niter) <= 1
        n = min(len(hist), len(ref[2]))
        assert np.allclose(hist[:n], ref[2][:n], rtol=1e-8)
        assert rel(x, ref[3]) < 1e-9
        assert np.array_equal(x, res[0][3]), "the solution must be bit-identical on every rank"
        assert np.array_equal(hist, res[0][2]), "all ranks must see the same scalars"
    info = [mats[0].shard_info(r) for r in range(nranks)]
    assert info[0]["row_begin"] == 0 and info[-1]["row_end"] == A.shape[0]
    assert all(info[r]["row_end"] == info[r + 1]["row_begin"] for r in range(nranks - 1))


@pytest.mark.parametrize("orth", [lib.ORTH_MGS, lib.ORTH_CGS2, lib.ORTH_CGS2_FUSED])
@pytest.mark.parametrize("nranks", [2, 4])
def test_gmres_sharded_short_run_parity(ctx, nranks, orth):
    """2-D inversion, fixed itmax: iteration counts and residual histories against the oracle and
    the single-rank kernel."""
    _, ops = workload("bowl_mixing", dim=2)
    A = ops["A"].tocsr()
    rng = np.random.default_rng(2)
    b = rng.uniform(-1, 1, A.shape[0])
    ps = ops["pscale"]

    def solve(c, dA):
        x = c.vector(A.shape[0])
        st, hist = lib.gmres_solve(dA, c.vector(b), x, pscale=ps, atol=1e-6, rtol=1e-6, itmax=130,
                                   memory=20, orth=orth, history=4096)
        return st.niter, bool(st.solved), hist, x.download()

    ref = solve(ctx, ctx.csr(A, drop_zeros=True))
    res, _, _ = _sharded(nranks, A, solve, drop_zeros=True)
    xo, so = krylov.gmres(A, b, x0=np.zeros(b.size), M=ps, atol=1e-6, rtol=1e-6, itmax=130, memory=20)
    for niter, solved, hist, x in res:
        assert niter == ref[0] and abs(niter - so.niter) <= 1
        n = min(len(hist), len(ref[2]), 100)
        assert np.allclose(hist[:n], ref[2][:n], rtol=1e-6)
        assert rel(x, ref[3]) < 1e-6
        assert np.array_equal(x, res[0][3])


def test_gmres_sharded_tight_matches_direct(ctx):
    """3-D evolution-sized SPD system solved by sharded GMRES to 1e-12: residual <= 1e-10, solution
    within 1e-8 of SuperLU; consecutive solves on one communicator (sequence numbers carry over)."""
    _, ops = workload("bowl_mixing")
    A = _evol(ops)
    rng = np.random.default_rng(3)
    bs = [rng.uniform(-1, 1, A.shape[0]) for _ in range(3)]
    dinv = 1.0 / A.diagonal()
    lu = spla.splu(A.tocsc())

    def solve(c, dA):
        out = []
        x = c.vector(A.shape[0])
        for b in bs:                              # warm-started from the previous answer
            st, _ = lib.gmres_solve(dA, c.vector(b), x, dinv=c.vector(dinv), atol=0.0, rtol=1e-12,
                                    memory=20, orth=lib.ORTH_CGS2)
            out.append((bool(st.solved), x.download()))
        return out

    res, _, _ = _sharded(2, A, solve)
    for rank_out in res:
        for (solved, x), b in zip(rank_out, bs):
            assert solved
            assert np.linalg.norm(A @ x - b) / np.linalg.norm(b) < 1e-10
            assert rel(x, lu.solve(b)) < 1e-8


def test_inversion_3d_sharded_first_iterations(ctx):
    """3-D inversion matrix (SM-resident form, RCM reorder, dropped zeros) on 2 ranks: history equal
    to the single-rank kernel over a bounded run."""
    _, ops = workload("bowl_mixing")
    A = ops["A"].tocsr()
    y = ops["B"] @ ops["b_init"] + ops["b0"] if "b_init" in ops else None
    rng = np.random.default_rng(4)
    b = rng.uniform(-1, 1, A.shape[0]) if y is None or not np.any(y) else y
    ps = ops["pscale"]

    def solve(c, dA):
        x = c.vector(A.shape[0])
        st, hist = lib.gmres_solve(dA, c.vector(b), x, pscale=ps, atol=1e-6, rtol=1e-6, itmax=400,
                                   memory=20, orth=lib.ORTH_CGS2, history=4096)
        return st.niter, hist, x.download(), st.device_ms

    ref = solve(ctx, ctx.csr(A, drop_zeros=True))
    res, _, _ = _sharded(2, A, solve, drop_zeros=True)
    for niter, hist, x, ms in res:
        assert niter == ref[0]
        assert np.allclose(hist[:300], ref[1][:300], rtol=1e-6)
        assert rel(x, ref[2]) < 1e-6


def test_inversion_3d_sharded_streaming_form(ctx):
    """Same system through the STREAMING form (forced: small footprint cap, several tiles per CTA) on 2
    ranks: the halo unpack feeds the comm warps' footprint gathers; fused CGS2 and MGS."""
    _, ops = workload("bowl_mixing")
    A = ops["A"].tocsr()
    b = np.random.default_rng(4).uniform(-1, 1, A.shape[0])
    ps = ops["pscale"]
    os.environ["NUPGCM_RESIDENT"] = "0"
    try:
        for orth in (lib.ORTH_CGS2_FUSED, lib.ORTH_MGS):
            def solve(c, dA):
                x = c.vector(A.shape[0])
                st, hist = lib.gmres_solve(dA, c.vector(b), x, pscale=ps, atol=1e-6, rtol=1e-6, itmax=130,
                                           memory=20, orth=orth, history=4096)
                return st.niter, hist, x.download()

            ref = solve(ctx, ctx.csr(A, drop_zeros=True))
            res, _, _ = _sharded(2, A, solve, drop_zeros=True)
            for niter, hist, x in res:
                assert niter == ref[0]
                assert np.allclose(hist[:100], ref[1][:100], rtol=1e-6)
                assert rel(x, ref[2]) < 1e-6
    finally:
        os.environ.pop("NUPGCM_RESIDENT")


@pytest.mark.parametrize("orth", [lib.ORTH_MGS, lib.ORTH_CGS2_FUSED])
def test_sharded_streaming_vector_forms_with_many_rows_per_cta(ctx, orth):
    """Once-refined bowl (N ~ 1.3e5) on 2 ranks of 24 CTAs: 2 700 rows per CTA, more than the register-resident
    Arnoldi forms of the sharded kernels hold (4 rows per thread), so the STREAMING Gram-Schmidt forms run with
    halo pushes — MGS with the new vector parked in the idle footprint arena (the path h = 0.02 takes on 8 GPUs)."""
    from nupgcm_b200 import workloads as W
    ops = W.host_operands(W.bowl_example(mesh=W.refined_bowl(1, h0=0.1)))
    A = ops["A"].tocsr()
    y = ops["B"] @ ops["b_init"] + ops["b0"]
    ps = ops["pscale"]

    def solve(c, dA):
        x = c.vector(A.shape[0])
        st, hist = lib.gmres_solve(dA, c.vector(y), x, pscale=ps, atol=0.0, rtol=1e-30, itmax=45,
                                   memory=20, orth=orth, history=64)
        return st.niter, hist, x.download()

    os.environ["NUPGCM_GRID"] = "48"           # single rank: the same 48 row blocks
    try:
        ref = solve(ctx, ctx.csr(A, drop_zeros=True))
    finally:
        os.environ.pop("NUPGCM_GRID")
    res, _, _ = _sharded(2, A, solve, drop_zeros=True, grid=24)
    for niter, hist, x in res:
        assert niter == ref[0] == 45
        assert np.allclose(hist, ref[1], rtol=1e-8)
        assert rel(x, ref[2]) < 1e-8


def test_model_steps_sharded_match_single_rank(ctx):
    """Three timesteps of the 2-D bowl with both solves sharded over 2 ranks (state replicated,
    element RHS replicated) against the single-GPU model."""
    import nupgcm_b200 as npg
    w, ops = workload("bowl_mixing", dim=2)

    def make(arch):
        inv = npg.InversionToolkit(arch, ops["A"], ops["pscale"], ops["B"], ops["b0"])
        ts = w.timestepper()
        evo = npg.EvolutionToolkit(arch, ops, w.params, w.forcings, ts)
        m = npg.Model(arch, w.params, w.forcings, w.fe_data(), inv, evo, ts, tables=ops["tables"])
        m.xb.upload(ops["b_init"])
        return m

    single = make(npg.GPU(0))
    npg.run_(single, n_steps=3)
    comms = local_ranks(2, ops["A"].shape[0])
    models = [make(npg.GPU(0, comm=c)) for c in comms]
    run_collective([lambda m=m: npg.run_(m, n_steps=3) for m in models])
    for m in models:
        for a, b in zip(m.step_log, single.step_log):
            assert abs(a["gmres_iters"] - b["gmres_iters"]) <= 1 and abs(a["cg_iters"] - b["cg_iters"]) <= 1
        assert rel(m.xb.download(), single.xb.download()) < 1e-8
        assert rel(m.inversion.solver.x.download(), single.inversion.solver.x.download()) < 1e-6
    assert np.array_equal(models[0].xb.download(), models[1].xb.download())


def test_channel_production_configuration_sharded_matches_single_rank(ctx):
    """BASELINE configs[3] on two ranks: periodic channel + basin box, P1 buoyancy, wind + surface flux, adaptive
    BDF1 (CFL from the device every step), the convection rebuild of Kv every step and the eddy-viscosity
    rebuild of the inversion matrix after step 10 — with BOTH solves sharded.  The rebuild kernels write the
    caller-order values on every rank; the sharded solver tables are refreshed from them before the next solve."""
    import nupgcm_b200 as npg
    from nupgcm_b200 import workloads as W
    w = W.with_b_order(W.channel_basin_box(periodic=True), 1)
    ops = W.host_operands(w)
    n = 12

    def make(arch):
        inv = npg.InversionToolkit(arch, ops["A"], ops["pscale"], ops["B"], ops["b0"], atol=0.0, rtol=1e-13,
                                   itmax=3000000, drop_zeros=False)
        ts = w.timestepper()
        evo = npg.EvolutionToolkit(arch, ops, w.params, w.forcings, ts, atol=0.0, rtol=1e-14)
        m = npg.Model(arch, w.params, w.forcings, w.fe_data(), inv, evo, ts, tables=ops["tables"])
        m.xb.upload(ops["b_init"])
        return m, ts

    def steps(m, ts):
        dts = []
        for _ in range(n):
            npg.run_(m, n_steps=1, resume=True)
            dts.append(ts.Δt)
        return dts

    single, ts1 = make(npg.GPU(0))
    dts1 = steps(single, ts1)
    comms = local_ranks(2, ops["A"].shape[0])
    models = [make(npg.GPU(0, comm=c)) for c in comms]
    out = run_collective([lambda m=m, t=t: steps(m, t) for m, t in models])
    for (m, _), dts in zip(models, out):
        assert np.allclose(dts, dts1, rtol=1e-9)
        assert rel(m.xb.download(), single.xb.download()) < 1e-8
        assert rel(m.inversion.solver.x.download(), single.inversion.solver.x.download()) < 1e-8
        assert all(r["gmres_solved"] and r["cg_solved"] for r in m.step_log)
    assert np.array_equal(models[0][0].xb.download(), models[1][0].xb.download())


def test_two_processes_over_ipc():
    """One process per GPU (torchrun, world size 2): CUDA IPC arenas, NVLink pushes."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29531",
           os.path.join(ROOT, "tools", "sharded_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "sharded_check ok" in out.stdout
