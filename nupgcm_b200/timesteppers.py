"""Timesteppers.  Mirrors reference ``src/timesteppers.jl:7-29`` (BDF1), ``:36-63`` (BDF2),
``:80-83`` (``update_t!``) and ``:108-122`` (``update_Δt!``: CFL-adaptive Δt for BDF1, evaluated
on the device by ``nupgcm_cfl_dt``; a no-op for BDF2)."""
from __future__ import annotations


class AbstractTimestepper:
    scheme = 0

    def __init__(self, t_start, t_stop, Δt, t=None):
        self.t_start = float(t_start)
        self.t = float(t_start if t is None else t)
        self.t_stop = float(t_stop)
        self.Δt = float(Δt)
        self.adaptive = False


class BDF1(AbstractTimestepper):
    scheme = 1

    def __init__(self, *, t_start, t_stop, Δt, t=None, adaptive=False, CFL_factor=0.8):
        super().__init__(t_start, t_stop, Δt, t)
        self.adaptive = bool(adaptive)
        self.CFL_factor = float(CFL_factor)


class BDF2(AbstractTimestepper):
    scheme = 2

    def __init__(self, *, t_start, t_stop, Δt, t=None):
        super().__init__(t_start, t_stop, Δt, t)


def update_t_(ts: AbstractTimestepper):
    """``update_t!`` (timesteppers.jl:80-83)."""
    ts.t += ts.Δt
    return ts


def update_Δt_(ts, mesh=None, u=None, u_min=0.01):
    """``update_Δt!`` (timesteppers.jl:108-122): Δt = CFL_factor · min_K h_K / max(|u|_{L∞(K)}, u_min)
    for an adaptive BDF1 — one kernel over the cells of ``mesh`` (a ``lib.ElementMesh``) and the
    device velocity ``u`` — and a no-op otherwise.

    DELIBERATE DEVIATION (pinned by tests/test_host_logic.py): the reference's ``update_Δt!(::BDF1, …)``
    has no ``adaptive`` check and ``run!`` calls it unconditionally (model.jl:131), so upstream a
    ``BDF1(adaptive=false)`` run also changes Δt every step — while its left-hand side is only re-formed
    when ``adaptive`` or the convection parameterisation is on (model.jl:251), i.e. the matrix and the
    right-hand side then disagree about Δt.  Here Δt changes only for ``adaptive=True``, which is the only
    configuration the reference exercises with BDF1 (scratch/run.jl:163)."""
    if getattr(ts, "adaptive", False):
        ts.Δt = mesh.cfl_dt(u, ts.CFL_factor, u_min)
    return ts


def evolution_parameter(params, ts) -> float:
    """θ of ``A = M + θ (Kₕ + Kᵥ)`` (evolution.jl:187-193)."""
    θ = ts.Δt * params.α ** 2 * params.ε ** 2 / params.μϱ
    return θ if ts.scheme == 1 else 2 / 3 * θ
