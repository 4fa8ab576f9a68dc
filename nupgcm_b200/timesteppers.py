"""Timesteppers.  Mirrors reference ``src/timesteppers.jl:7-29`` (BDF1), ``:36-63`` (BDF2),
``:80-83`` (``update_t!``).  ``update_Δt!`` (adaptive CFL, ``:108-119``) is a "next" row of the
scope table: ``adaptive=True`` is rejected for now."""
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
        if adaptive:
            raise NotImplementedError("adaptive Δt is not on the B200 path yet (SURVEY.md §8 f-1)")
        self.CFL_factor = CFL_factor


class BDF2(AbstractTimestepper):
    scheme = 2

    def __init__(self, *, t_start, t_stop, Δt, t=None):
        super().__init__(t_start, t_stop, Δt, t)


def update_t_(ts: AbstractTimestepper):
    """``update_t!`` (timesteppers.jl:80-83)."""
    ts.t += ts.Δt
    return ts


def update_Δt_(ts, *args, **kw):
    """``update_Δt!`` — a no-op for fixed-Δt steppers (timesteppers.jl:120-122)."""
    return ts


def evolution_parameter(params, ts) -> float:
    """θ of ``A = M + θ (Kₕ + Kᵥ)`` (evolution.jl:187-193)."""
    θ = ts.Δt * params.α ** 2 * params.ε ** 2 / params.μϱ
    return θ if ts.scheme == 1 else 2 / 3 * θ
