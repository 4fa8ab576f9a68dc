"""CPU restatement of nuPGCM's time loop.  TEST INFRASTRUCTURE ONLY (never imported by the
product; see oracle/krylov.py for the rule).

Restates ``run!`` (reference ``src/model.jl:90-211``), ``evolve!`` (``:213-285``), ``invert!``
(``src/inversion.jl:101-110``) and the LHS switch of ``src/evolution.jl:110-111,133-177``,
including the quirks of SURVEY.md App. D: ``while t < t_stop`` with ``t += Δt`` (51 steps for
``t_stop = 50Δt``), the hybrid first BDF2 step (BDF1 left-hand side, BDF2 right-hand side with
``prev = curr``) and the LHS switch at ``i == 2``.

Two solver modes:
* ``"direct"`` — what the reference's CPU path does: LU factor once, solve per step
  (``src/iterative_solvers.jl:42-48`` with ``lu(A)`` from ``src/inversion.jl:58`` and
  ``src/evolution.jl:152,170``; SciPy SuperLU stands in for UMFPACK);
* ``"krylov"`` — the GPU path's algorithm on the CPU: warm-started GMRES(20)/CG of
  ``oracle/krylov.py`` with the reference's default tolerances; gives the iteration counts the
  CUDA path must reproduce.

Operands come from ``nupgcm_b200.workloads.host_operands`` (host-side set-up, pinned to the
reference's matrix fixture by tests/test_fe_setup.py); everything on the per-step path
(element RHS, combine, SpMV, solves) is restated here independently of the product.
"""
from __future__ import annotations

import time

import numpy as np
import scipy.sparse.linalg as spla

from . import krylov
from .element_rhs import cfl_dt, kv_rebuild, nu_friction, rhs_adv, rhs_combine


class CpuModel:
    def __init__(self, ops: dict, params: dict, scheme: int, dt: float, t_start: float,
                 t_stop: float, solver: str = "direct", atol=1e-6, rtol=1e-6, memory=20,
                 orth="mgs", adaptive=False, cfl_factor=0.8, conv=None, kv_q=None, eddy=None,
                 f_q=None, A0=None):
        self.ops = ops
        self.α, self.ε, self.μϱ, self.N2 = (params[k] for k in ("α", "ε", "μϱ", "N2"))
        self.scheme, self.dt, self.t, self.t_stop = scheme, dt, t_start, t_stop
        self.solver = solver
        # BDF1(adaptive=true): Δt from the CFL condition every step (timesteppers.jl:108-119) and the
        # LHS re-formed (and, on the reference's CPU path, re-factorised) every step (model.jl:251-261)
        self.adaptive, self.cfl_factor = adaptive, cfl_factor
        self.dts = []
        # ConvectionParameterization(κᶜ, N²min): Kᵥ, rhsᵥ, rhs_diff rebuilt from b every step
        # (model.jl:229-246); `kv_q` = base κᵥ at the quadrature points
        self.conv, self.kv_q = conv, kv_q
        # EddyParameterization(f, N²min): eddy = N²min, f_q = f at the quadrature points, A0 = the
        # frictionless part of the inversion matrix; A rebuilt every 10 steps (model.jl:160-170)
        self.eddy, self.f_q, self.A0 = eddy, f_q, A0
        if conv is not None or eddy is not None:
            self.ops = dict(ops)
        self.kw = dict(atol=atol, rtol=rtol)
        self.memory, self.orth = memory, orth
        self.xu = np.zeros(ops["A"].shape[0])        # [u; p] in solver order
        self.xb = ops["b_init"].copy()
        self.i = 1
        self.log = []
        self._lu_A = None
        self._lhs(1)                                  # BDF1 LHS first (evolution.jl:110-111)

    def theta(self, scheme):
        θ = self.dt * self.α ** 2 * self.ε ** 2 / self.μϱ
        return θ if scheme == 1 else 2.0 / 3.0 * θ

    def _lhs(self, scheme):
        o = self.ops
        self.A_evol = (o["M"] + self.theta(scheme) * (o["Kh"] + o["Kv"])).tocsr()
        if self.solver == "direct":
            self._lu_evol = spla.splu(self.A_evol.tocsc())
        else:
            self.dinv = 1.0 / self.A_evol.diagonal()

    # -- invert! --------------------------------------------------------------------------
    def invert(self):
        o = self.ops
        y = o["B"] @ self.xb + o["b0"]
        if self.solver == "direct":
            if self._lu_A is None:
                self._lu_A = spla.splu(o["A"].tocsc())
            self.xu = self._lu_A.solve(y)
            return 0
        M = np.full(y.size, o["pscale"])
        self.xu, st = krylov.gmres(o["A"], y, x0=self.xu, M=M, memory=self.memory, orth=self.orth,
                                   history=False, **self.kw)
        return st.niter

    # -- evolve! --------------------------------------------------------------------------
    def evolve(self, u_prev, b_prev):
        o = self.ops
        nu = o["nu"]
        θ = self.theta(self.scheme)
        if self.conv is not None:
            o["Kv"], o["rhs_v"], o["rhs_diff"] = kv_rebuild(o["tables"], self.kv_q, self.α, self.N2,
                                                             self.conv[0], self.conv[1], self.xb, o["M"])
        if self.adaptive or self.conv is not None:
            self._lhs(self.scheme)      # θ of the current timestepper, also on a BDF2 run's first step (model.jl:227,251-255)
        adv = rhs_adv(o["tables"], self.scheme, self.dt, self.N2, self.xb, b_prev,
                      self.xu[:nu], u_prev[:nu])
        y = rhs_combine(adv, θ, self.dt, o["rhs_diff"], o["rhs_flux"], o["rhs_m"], o["rhs_h"],
                        o["rhs_v"])
        if self.solver == "direct":
            self.xb = self._lu_evol.solve(y)
            return 0
        self.xb, st = krylov.cg(self.A_evol, y, x0=self.xb, M=self.dinv, history=False, **self.kw)
        return st.niter

    # -- run! -----------------------------------------------------------------------------
    def run(self, n_steps=None):
        u_prev, b_prev = self.xu.copy(), self.xb.copy()
        done = 0
        while self.t < self.t_stop and (n_steps is None or done < n_steps):
            if self.adaptive:
                self.dt = cfl_dt(self.ops["tables"], self.xu[:self.ops["nu"]], self.cfl_factor)
                self.dts.append(self.dt)
            if self.i == 2 and self.scheme == 2:
                self._lhs(2)
            u_curr, b_curr = self.xu.copy(), self.xb.copy()
            t0 = time.perf_counter()
            cg_it = self.evolve(u_prev, b_prev)
            gm_it = self.invert()
            self.t += self.dt
            u_max = np.abs(self.xu[:self.ops["nu"]]).max()
            b_max = np.abs(self.xb).max()
            if max(u_max, b_max) > 1e3 or np.isnan(u_max) or np.isnan(b_max):
                raise RuntimeError("Blow-up detected, stopping simulation")
            u_prev, b_prev = u_curr, b_curr
            if self.eddy is not None and self.i % 10 == 0:
                o = self.ops
                o["A"] = (self.A0 + nu_friction(o["tables"], self.f_q, self.α ** 2 * self.ε ** 2, self.α,
                                                self.N2, self.eddy, self.xb, self.A0.shape[0])).tocsr()
                self._lu_A = None
            self.log.append({"i": self.i, "cg_iters": cg_it, "gmres_iters": gm_it,
                             "seconds": time.perf_counter() - t0})
            self.i += 1
            done += 1
        return self


def cpu_model_for(workload, ops=None, **kw):
    """Build the CPU oracle for a ``nupgcm_b200.workloads.Workload``."""
    from nupgcm_b200.workloads import host_operands
    ops = host_operands(workload) if ops is None else ops
    p = workload.params
    tk = workload.timestepper_kwargs
    scheme = kw.pop("scheme", 1 if tk.get("adaptive") else 2)
    if tk.get("adaptive"):
        kw.setdefault("adaptive", True)
        kw.setdefault("cfl_factor", tk.get("CFL_factor", 0.8))
    f = workload.forcings
    fe = workload.fe_data()
    if f.conv_param.is_on and "conv" not in kw:
        kw["conv"] = (f.conv_param.κᶜ, f.conv_param.N2min)
        kw["kv_q"] = fe.mesh.dΩ.coefficient(f.κᵥ, slice(None))
    if f.eddy_param.is_on and "eddy" not in kw:
        from nupgcm_b200._forms import build_A_inversion      # host set-up code (operands), not the device path
        pi = fe.dofs.p_inversion
        kw["eddy"] = f.eddy_param.N2min
        kw["f_q"] = fe.mesh.dΩ.coefficient(f.eddy_param.f, slice(None))
        kw["A0"] = build_A_inversion(fe, p, 0.0)[pi][:, pi].tocsr()
    m = CpuModel(ops, {"α": p.α, "ε": p.ε, "μϱ": p.μϱ, "N2": p.N2}, scheme, tk["Δt"], tk["t_start"],
                 tk["t_stop"], **kw)
    return m
