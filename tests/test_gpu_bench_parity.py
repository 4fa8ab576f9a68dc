"""Parity of exactly what bench.py times ("close the parity holes on what you benchmark"):

* the BASELINE configs[1] set-up (shipped bowl3D h = 0.08 mesh, examples/bowl_mixing.jl) stepped at
  tight tolerance with BOTH orthogonalisations the benchmark reports — `mgs` (Krylov.jl's, the API
  default) and `cgs2f` (two-reduction CGS2) — against the oracle's direct-solve path (the reference's
  CPU algorithm): u, b within 1e-8 relative L2 after N steps, solver relative residual <= 1e-10;
* the same with both solves row-block sharded over 2 and 4 ranks;
* the headline mesh (h = 0.04, N = 263 159, streamed-matrix kernels): the initial inversion driven to
  convergence, its true residual checked on the host with SciPy.  A SuperLU comparison is not possible
  there: SciPy's factorisation of that matrix was stopped after 13 minutes and 7.6 GB in this container."""
import numpy as np
import pytest

from conftest import workload
import nupgcm_b200 as npg
from nupgcm_b200 import lib
from nupgcm_b200 import workloads as W
from nupgcm_b200.sharding import local_ranks, run_collective
from oracle.stepping import cpu_model_for

pytestmark = pytest.mark.gpu

ORTH = {"mgs": lib.ORTH_MGS, "cgs2f": lib.ORTH_CGS2_FUSED}
N_STEPS = 3
TIGHT = dict(atol=0.0, rtol=1e-12, itmax=4000000)
_CACHE = {}


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def example():
    if "ex" not in _CACHE:
        w, ops = workload("bowl_example")
        cpu = cpu_model_for(w, ops, solver="direct")
        cpu.invert()                                   # examples/bowl_mixing.jl:194
        cpu.run(n_steps=N_STEPS)
        _CACHE["ex"] = (w, ops, cpu)
    return _CACHE["ex"]


def make(arch, w, ops, orth):
    inv = npg.InversionToolkit(arch, ops["A"], ops["pscale"], ops["B"], ops["b0"], orth=ORTH[orth], history=False, **TIGHT)
    ts = w.timestepper()
    evo = npg.EvolutionToolkit(arch, ops, w.params, w.forcings, ts, atol=0.0, rtol=1e-14, history=False)
    m = npg.Model(arch, w.params, w.forcings, w.fe_data(), inv, evo, ts, tables=ops["tables"])
    m.xb.upload(ops["b_init"])
    return m


def check_against_direct(m, ops, cpu, nu):
    xu, xb = m.inversion.solver.x.download(), m.xb.download()
    assert rel(xb, cpu.xb) < 1e-8
    assert rel(xu[:nu], cpu.xu[:nu]) < 1e-8
    y = ops["B"] @ xb + ops["b0"]
    assert np.linalg.norm(y - ops["A"] @ xu) / np.linalg.norm(y) < 1e-10
    assert all(r["gmres_solved"] and r["cg_solved"] for r in m.step_log)


@pytest.mark.parametrize("orth", ["mgs", "cgs2f"])
def test_bench_secondary_config_fields_match_direct_oracle(orth):
    w, ops, cpu = example()
    m = make(npg.GPU(0), w, ops, orth)
    npg.invert_(m)
    npg.run_(m, n_steps=N_STEPS)
    check_against_direct(m, ops, cpu, w.fe_data().dofs.nu)


@pytest.mark.parametrize("nranks,orth", [(2, "mgs"), (2, "cgs2f"), (4, "cgs2f")])
def test_bench_secondary_config_sharded_fields_match_direct_oracle(nranks, orth):
    w, ops, cpu = example()
    comms = local_ranks(nranks, ops["A"].shape[0])
    models = [make(npg.GPU(0, comm=c), w, ops, orth) for c in comms]

    def go(m):
        npg.invert_(m)
        npg.run_(m, n_steps=N_STEPS)
    run_collective([lambda m=m: go(m) for m in models])
    for m in models:
        check_against_direct(m, ops, cpu, w.fe_data().dofs.nu)
    assert np.array_equal(models[0].xb.download(), models[-1].xb.download())


def test_headline_mesh_initial_inversion_converges_and_residual_checks_on_host():
    """h = 0.04 (bench.py's `value`): cold-start GMRES(20) to the reference tolerance through the streamed-matrix
    kernels; the stopping test is ‖M r‖ <= atol + rtol ‖M r₀‖ with M = I/h³ (src/inversion.jl:54), i.e.
    ‖r‖ <= h³ atol + rtol ‖y‖ from x₀ = 0 — verified here with SciPy's product, not the device's."""
    w = W.bowl_example(mesh=W.refined_bowl(1))
    fe = w.fe_data()
    from nupgcm_b200.inversion import permuted_inversion_system
    A, B, b0, pscale = permuted_inversion_system(fe, w.params, w.forcings)
    y = B @ fe.spaces.B.interpolate(w.b0)[0][fe.dofs.p_b] + b0
    ctx = npg.GPU(0).ctx
    dA = ctx.csr(A, drop_zeros=True)
    out = {}
    for name in ("cgs2f", "mgs"):
        x = ctx.vector(y.size)
        st, _ = lib.gmres_solve(dA, ctx.vector(y), x, pscale=pscale, atol=1e-6, rtol=1e-6, memory=20, orth=ORTH[name])
        xh = x.download()
        r = np.linalg.norm(y - A @ xh)
        assert st.solved
        assert r <= 1.5 * (1e-6 / pscale + 1e-6 * np.linalg.norm(y)), (name, r)
        out[name] = (st.niter, xh)
    # the two orthogonalisations: same answer to the solver tolerance, iteration counts inside the spread of
    # long restarted runs
    assert abs(out["mgs"][0] - out["cgs2f"][0]) <= 0.10 * out["mgs"][0]
    assert rel(out["cgs2f"][1], out["mgs"][1]) < 1e-4
