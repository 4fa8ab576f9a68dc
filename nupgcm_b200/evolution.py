"""``EvolutionToolkit``.  Mirrors reference ``src/evolution.jl:1-17`` (struct), ``:55-131``
(set-up), ``:133-177`` (``collect_evolution_LHS[!]``) and ``:187-193`` (θ)."""
from __future__ import annotations

import numpy as np

from . import lib
from ._forms import build_Kh, build_Kv, build_M, build_rhs_diff, build_rhs_flux
from .architectures import GPU, on_architecture
from .dofs import FEData
from .inputs import Forcings, Parameters
from .iterative_solvers import Diagonal, IterativeSolverToolkit
from .timesteppers import BDF1, AbstractTimestepper, evolution_parameter


def permuted_evolution_system(fe_data: FEData, params: Parameters, forcings: Forcings):
    """Host operands in solver order (evolution.jl:80-99): M, Kₕ, Kᵥ (identical sparsity
    patterns) and the five RHS vectors."""
    p = fe_data.dofs.p_b
    M, rhs_m = build_M(fe_data)
    Kh, rhs_h = build_Kh(fe_data, forcings.κₕ)
    Kv, rhs_v = build_Kv(fe_data, forcings.κᵥ)
    out = {}
    for name, mat in (("M", M), ("Kh", Kh), ("Kv", Kv)):
        m = mat[p][:, p].tocsr()
        m.sort_indices()
        out[name] = m
    if not (np.array_equal(out["M"].indices, out["Kh"].indices) and
            np.array_equal(out["M"].indices, out["Kv"].indices)):
        raise AssertionError("M, Kh, Kv must share a sparsity pattern")
    out["rhs_diff"] = build_rhs_diff(params, fe_data, forcings.κᵥ)[p]
    out["rhs_flux"] = build_rhs_flux(params, forcings, fe_data)[p]
    out["rhs_m"], out["rhs_h"], out["rhs_v"] = rhs_m[p], rhs_h[p], rhs_v[p]
    return out


class EvolutionToolkit:
    def __init__(self, arch, fe_data_or_ops, params: Parameters, forcings: Forcings,
                 ts: AbstractTimestepper, atol=1e-6, rtol=1e-6, itmax=0, history=True,
                 verbose=False):
        if not isinstance(arch, GPU):
            raise NotImplementedError("nupgcm_b200 only provides the GPU() architecture; the CPU "
                                      "path is the reference's own (no fallback)")
        ops = (fe_data_or_ops if isinstance(fe_data_or_ops, dict)
               else permuted_evolution_system(fe_data_or_ops, params, forcings))
        self.arch = arch
        self.M = on_architecture(arch, ops["M"])
        self.Kh = on_architecture(arch, ops["Kh"])
        self.Kv = on_architecture(arch, ops["Kv"])
        for k in ("rhs_diff", "rhs_flux", "rhs_m", "rhs_h", "rhs_v"):
            setattr(self, k, on_architecture(arch, ops[k]))
        nb = ops["M"].shape[0]
        # LHS storage with the shared pattern; values filled by collect_evolution_LHS_
        self._A = on_architecture(arch, ops["M"])
        if getattr(arch, "comm", None) is not None:
            self._A.shard(arch.comm)                            # CG becomes collective over the ranks
        self._dinv = arch.ctx.vector(nb)
        y = arch.ctx.vector(nb)
        x = arch.ctx.vector(nb)
        kwargs = dict(atol=atol, rtol=rtol, itmax=itmax, history=history, verbose=verbose)
        self.solver = IterativeSolverToolkit(self._A, Diagonal(self._dinv), x, y, "cg", kwargs,
                                             "Evolution")
        # always start with a BDF1 left-hand side (evolution.jl:110-111)
        ts1 = BDF1(t_start=ts.t_start, t_stop=ts.t_stop, Δt=ts.Δt)
        collect_evolution_LHS_(self, params, forcings, ts1)


def collect_evolution_LHS_(evolution: EvolutionToolkit, params, forcings, ts):
    """``collect_evolution_LHS!`` (evolution.jl:133-177): A = M + θ(Kₕ+Kᵥ), P = 1/diag(A) —
    one value-combine kernel and one diagonal kernel on the device instead of a host sparse add,
    CSC->CSR conversion and upload."""
    θ = evolution_parameter(params, ts)
    evolution._A.combine(evolution.M, evolution.Kh, evolution.Kv, θ)
    evolution._A.inv_diag(evolution._dinv)
    evolution.solver.A = evolution._A
    evolution.solver.P = Diagonal(evolution._dinv)
    return evolution
