"""``State``, ``Model``, ``run!``, ``evolve!``, ``invert!``, ``sync_flow!``.

Mirrors reference ``src/model.jl:1-28`` (structs), ``:47-75`` (constructors), ``:77-88``
(``set_b!``), ``:90-211`` (``run!``), ``:213-285`` (``evolve!``), ``:302-317``
(``invert!`` / ``sync_flow!``).  Julia's ``f!`` is spelled ``f_`` here.

Difference by design: the state lives on the device in solver order between steps.  The
reference keeps it on the host in Gridap order and crosses PCIe four times per step
(model.jl:275,282,312; inversion.jl:104); here host copies are made only on request
(``State.u/p/b`` properties, or ``run_(..., sync_state=True)`` which reproduces the reference's
per-step host synchronisation for end-to-end timing)."""
from __future__ import annotations

import time

import numpy as np

from . import lib
from .architectures import GPU
from .dofs import FEData
from .element_tables import element_tables
from .evolution import EvolutionToolkit, collect_evolution_LHS_
from .inputs import Forcings, Parameters
from .inversion import InversionToolkit, invert_ as _invert_toolkit
from .iterative_solvers import iterative_solve_
from .timesteppers import AbstractTimestepper, BDF2, evolution_parameter, update_t_, update_Δt_


class State:
    """u, p, b.  Device-resident: ``xu`` = [u; p] and ``xb`` = b in solver order; the
    properties give host copies in Gridap order (what ``save_state`` stores, IO.jl:1-10)."""

    def __init__(self, model: "Model"):
        self._m = model

    @property
    def u(self):
        m = self._m
        x = m.inversion.solver.x.download()
        return x[m.fe_data.dofs.inv_p_inversion][:m.fe_data.dofs.nu]

    @property
    def p(self):
        m = self._m
        x = m.inversion.solver.x.download()
        return x[m.fe_data.dofs.inv_p_inversion][m.fe_data.dofs.nu:]

    @property
    def b(self):
        m = self._m
        return m.xb.download()[m.fe_data.dofs.inv_p_b]


class Model:
    def __init__(self, arch, params: Parameters, forcings: Forcings, fe_data: FEData,
                 inversion: InversionToolkit, evolution: EvolutionToolkit | None = None,
                 timestepper: AbstractTimestepper | None = None, tables: dict | None = None):
        if not isinstance(arch, GPU):
            raise NotImplementedError("nupgcm_b200 only provides the GPU() architecture")
        self.arch = arch
        self.params = params
        self.forcings = forcings
        self.fe_data = fe_data
        self.inversion = inversion
        self.evolution = evolution
        self.timestepper = timestepper
        ctx = arch.ctx
        nb = fe_data.dofs.nb
        N = fe_data.dofs.nu + fe_data.dofs.np
        # rest state (model.jl:64-75): solver vectors are zero-initialised
        self.xb = evolution.solver.x if evolution is not None else ctx.vector(nb)
        self.state = State(self)
        self.step_log = []
        if evolution is not None:
            self.mesh = lib.ElementMesh(ctx, tables if tables is not None else element_tables(fe_data))
            self._rhs_adv = ctx.vector(nb)
            if forcings.conv_param.is_on:
                # base κᵥ at the quadrature points of every cell; the device adds the convective part
                kv_q = fe_data.mesh.dΩ.coefficient(forcings.κᵥ, slice(None))
                self.mesh.enable_kv_rebuild(evolution.Kv, kv_q)
            if forcings.eddy_param.is_on:
                # the frictionless part of the inversion matrix stays on the device; the friction
                # block is re-assembled from ν(∂z b) every 10 steps (model.jl:160-170)
                from ._forms import build_A_inversion
                info = inversion.solver.A.info()
                if info["nnz_stored"] != info["nnz_given"]:
                    raise ValueError("the eddy parameterisation changes the value pattern of the "
                                     "inversion matrix: build InversionToolkit(..., drop_zeros=False)")
                p = fe_data.dofs.p_inversion
                A0 = build_A_inversion(fe_data, params, 0.0)[p][:, p].tocsr()
                A0.sort_indices()
                f_q = fe_data.mesh.dΩ.coefficient(forcings.eddy_param.f, slice(None))
                self.mesh.enable_nu_rebuild(inversion.solver.A, A0.data, f_q)
            self._b_prev, self._b_curr = ctx.vector(nb), ctx.vector(nb)
            self._u_prev, self._u_curr = ctx.vector(N), ctx.vector(N)


def set_b_(model: Model, b):
    """``set_b!`` (model.jl:77-88): a function of x (interpolated) or free values in Gridap order."""
    Bs = model.fe_data.spaces.B
    vals = Bs.interpolate(b)[0] if callable(b) else np.asarray(b, dtype=np.float64)
    model.xb.upload(vals[model.fe_data.dofs.p_b])
    return model


def invert_(model: Model, b: lib.Vector | None = None):
    """``invert!(model)`` (model.jl:302-309).  ``sync_flow!`` is implicit: the flow *is* the
    solver vector."""
    _invert_toolkit(model.inversion, model.xb if b is None else b)
    return model


def sync_flow_(model: Model):
    """``sync_flow!`` (model.jl:311-317): host copies of u and p in Gridap order."""
    x = model.inversion.solver.x.download()[model.fe_data.dofs.inv_p_inversion]
    nu = model.fe_data.dofs.nu
    return x[:nu], x[nu:]


def evolve_(model: Model, u_prev: lib.Vector, b_prev: lib.Vector):
    """``evolve!`` (model.jl:213-285): element RHS -> combine -> CG, all on the device."""
    ev = model.evolution
    ts = model.timestepper
    solver = ev.solver
    θ = evolution_parameter(model.params, ts)                      # model.jl:227
    conv = model.forcings.conv_param
    if conv.is_on:
        # κᵥ(α ∂z b_total) -> Kᵥ, rhsᵥ, rhs_diff re-assembled on the device (model.jl:229-246)
        model.mesh.rebuild_kv(model.params.α, model.params.N2, conv.κᶜ, conv.N2min, model.xb,
                              ev.Kv, ev.rhs_v, ev.rhs_diff)
    if conv.is_on or ts.adaptive:
        # A = M + θ(Kₕ+Kᵥ), P = 1/diag(A) with this step's Δt / Kᵥ (model.jl:251-261): two kernels on
        # the device-resident operands instead of a host sparse add + CSC->CSR + upload
        collect_evolution_LHS_(ev, model.params, model.forcings, ts)
    model.mesh.rhs_adv(ts.scheme, ts.Δt, model.params.N2, model.xb, b_prev,
                       model.inversion.solver.x, u_prev, model._rhs_adv)   # model.jl:269-275
    lib.rhs_combine(solver.y, model._rhs_adv, θ, ts.Δt, ev.rhs_diff, ev.rhs_flux, ev.rhs_m,
                    ev.rhs_h, ev.rhs_v)                             # model.jl:278
    iterative_solve_(solver)                                        # model.jl:279
    return model


def _sync_buffers(model: Model) -> dict:
    """Device index vectors and staging vectors of the host-synchronised mode (built once)."""
    sy = getattr(model, "_sync", None)
    if sy is None:
        ctx, d = model.arch.ctx, model.fe_data.dofs
        N = d.nu + d.np
        sy = {"p_b": ctx.index(d.p_b), "inv_p_b": ctx.index(d.inv_p_b),
              "p_inversion": ctx.index(d.p_inversion), "inv_p_inversion": ctx.index(d.inv_p_inversion),
              "stage_b": ctx.vector(d.nb), "stage_x": ctx.vector(N),
              # page-locked host copies: the per-step transfers are plain DMA
              "host_x": ctx.pinned_array(N), "host_b": ctx.pinned_array(d.nb)}
        model._sync = sy
    return sy


def run_(model: Model, n_info=10, n_save=float("inf"), n_steps=None, sync_state=False,
         host_state=None, log=None, advection=True, resume=False, out_dir=None, save_history=False):
    """``run!`` (model.jl:90-211).  ``n_steps`` bounds the number of steps taken by this call
    (the reference loops until ``t >= t_stop``); ``resume=True`` makes a call continue the previous
    one (see below).  With ``sync_state`` the state is copied to the
    host (Gridap order) and back every step, reproducing the reference's PCIe pattern."""
    ts = model.timestepper
    xu = model.inversion.solver.x
    xb = model.xb
    nu = model.fe_data.dofs.nu
    dofs = model.fe_data.dofs
    # copies of previous and current u, b (model.jl:120-123).  ``resume`` continues the run of an
    # earlier ``run_`` call on this model instead (keeps the previous-step fields, so that stepping
    # one step per call is the same computation as one call for all steps).
    if not (resume and getattr(model, "_has_prev", False)):
        model._u_prev.copy_from(xu)
        model._b_prev.copy_from(xb)
    model._has_prev = True
    i = getattr(model, "_step_index", 1)
    done = 0
    while ts.t < ts.t_stop and (n_steps is None or done < n_steps):
        update_Δt_(ts, model.mesh, xu)                              # model.jl:131
        if i == 2 and isinstance(ts, BDF2):
            collect_evolution_LHS_(model.evolution, model.params, model.forcings, ts)  # :134-137
        if sync_state and host_state is not None:
            # host -> device of the state the step starts from: one copy per field in Gridap order,
            # permuted to solver order by a gather kernel (x[p], model.jl:274; inversion.jl:37-39)
            sy = _sync_buffers(model)
            sy["host_b"][:] = host_state["b"]
            sy["stage_b"].upload(sy["host_b"])
            xb.gather_from(sy["stage_b"], sy["p_b"])
            sy["host_x"][:nu] = host_state["u"]
            sy["host_x"][nu:] = host_state["p"]
            sy["stage_x"].upload(sy["host_x"])
            xu.gather_from(sy["stage_x"], sy["p_inversion"])
        model._u_curr.copy_from(xu)                                 # model.jl:140-141
        model._b_curr.copy_from(xb)
        evolve_(model, model._u_prev, model._b_prev)                # model.jl:144
        invert_(model)                                              # model.jl:145
        update_t_(ts)                                               # model.jl:146
        u_max, u_nan = xu.maxabs(nu)                                # model.jl:149-153
        b_max, b_nan = xb.maxabs()
        if max(u_max, b_max) > 1e3 or u_nan or b_nan:
            raise RuntimeError("Blow-up detected, stopping simulation")
        model._u_prev, model._u_curr = model._u_curr, model._u_prev  # model.jl:156-157
        model._b_prev, model._b_curr = model._b_curr, model._b_prev
        eddy = model.forcings.eddy_param
        if eddy.is_on and advection and i % 10 == 0:                # model.jl:160-170
            p = model.params
            model.mesh.rebuild_friction(p.α ** 2 * p.ε ** 2, p.α, p.N2, eddy.N2min, 10.0, 1.0, xb,
                                          model.inversion.solver.A)
        if sync_state and host_state is not None:
            # device -> host in Gridap order (solver.x[inv_perm], model.jl:282,312): gather on the
            # device, one copy per field
            sy = _sync_buffers(model)
            sy["stage_x"].gather_from(xu, sy["inv_p_inversion"])
            x = sy["stage_x"].download(out=sy["host_x"])
            host_state["u"], host_state["p"] = x[:nu].copy(), x[nu:].copy()
            sy["stage_b"].gather_from(xb, sy["inv_p_b"])
            host_state["b"] = sy["stage_b"].download(out=sy["host_b"]).copy()
        rec = {"i": i, "t": ts.t, "cg_iters": model.evolution.solver.stats.niter,
               "gmres_iters": model.inversion.solver.stats.niter,
               "cg_ms": model.evolution.solver.stats.timer * 1e3,
               "gmres_ms": model.inversion.solver.stats.timer * 1e3,
               "cg_solved": model.evolution.solver.stats.solved,
               "gmres_solved": model.inversion.solver.stats.solved,
               "u_max": u_max, "b_max": b_max}
        model.step_log.append(rec)
        if log is not None and i % n_info == 0:
            log(rec)
        if n_save != float("inf") and i % int(n_save) == 0:         # model.jl:194-197 (save_vtk is out of scope)
            if out_dir is None:
                raise ValueError("run_(..., n_save=k) needs out_dir: the directory the state files go to "
                                 "(the reference writes $out_dir/data/state_%016d.jld2)")
            import os
            from .io import save_state
            os.makedirs(os.path.join(out_dir, "data"), exist_ok=True)
            model._step_index = i + 1
            save_state(model, os.path.join(out_dir, "data", "state_%016d.jld2" % i), history=save_history)
        i += 1
        done += 1
    model._step_index = i
    return model
