"""Configuration structs: ``Parameters``, surface boundary conditions, ``Forcings``.

Mirrors reference ``src/inputs.jl:3-15`` (Parameters), ``:33-59`` (surface BCs) and ``:141-189``
(Forcings).  Functions of space take an ``(n, 3)`` array of points and return ``(n,)`` values
(the vectorised form of the reference's ``x -> ...`` closures); plain numbers are accepted too.
``ConvectionParameterization`` (``inputs.jl:62-91``) is on the device path (per-step Kᵥ rebuild,
``nupgcm_rebuild_kv``); the eddy parameterisation (``inputs.jl:95-137``, friction block of the
inversion matrix every 10 steps) is the remaining half of SURVEY.md §8 f-2 and is rejected loudly.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Callable, Union

Coef = Union[float, Callable]


@dataclass
class Parameters:
    ε: float          # Ekman number
    α: float          # aspect ratio
    μϱ: float         # Prandtl times Burger number
    N2: float         # background stratification N²
    f: Coef           # Coriolis parameter f(x)
    H: Coef           # depth H(x)

    def __post_init__(self):
        self.ε, self.α, self.μϱ, self.N2 = (float(v) for v in (self.ε, self.α, self.μϱ, self.N2))


@dataclass
class SurfaceDirichletBC:
    value: Coef


@dataclass
class SurfaceFluxBC:
    flux: Coef


@dataclass
class _Off:
    is_on: bool = False


@dataclass
class ConvectionParameterization:
    """``ConvectionParameterization(; κᶜ, N²min)`` (inputs.jl:62-91):
    κᵥ ← κᵥ + κᶜ (1 + tanh(−α ∂z b_total / N²min)) / 2."""
    κᶜ: float
    N2min: float
    is_on: bool = True


def κᵥ_convection(conv_param: ConvectionParameterization, κᵥ, αbz):
    """inputs.jl:87-91, on arrays of quadrature-point values."""
    import numpy as np
    return κᵥ + conv_param.κᶜ * (1 + np.tanh(-αbz / conv_param.N2min)) / 2


@dataclass
class Forcings:
    ν: Coef
    κₕ: Coef
    κᵥ: Coef
    τˣ: Coef
    τʸ: Coef
    b_surface_bc: Any
    conv_param: Any = field(default_factory=_Off)
    eddy_param: Any = field(default_factory=_Off)

    def __post_init__(self):
        if self.eddy_param.is_on:
            raise NotImplementedError(
                "the eddy parameterisation (ν rebuild of the inversion matrix) is not on the B200 "
                "hot path yet (SURVEY.md §8 f-2)")
