"""Configuration structs: ``Parameters``, surface boundary conditions, ``Forcings``.

Mirrors reference ``src/inputs.jl:3-15`` (Parameters), ``:33-59`` (surface BCs) and ``:141-189``
(Forcings).  Functions of space take an ``(n, 3)`` array of points and return ``(n,)`` values
(the vectorised form of the reference's ``x -> ...`` closures); plain numbers are accepted too.
``ConvectionParameterization`` (``inputs.jl:62-91``; per-step Kᵥ rebuild, ``nupgcm_rebuild_kv``)
and ``EddyParameterization`` (``inputs.jl:95-137``; friction block of the inversion matrix every
10 steps, ``nupgcm_rebuild_friction``) are both evaluated on the device.
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
class EddyParameterization:
    """``EddyParameterization(; f, N²min)`` (inputs.jl:95-137): ν = f²/(α ∂z b_total), smoothly
    limited to ν_min ≤ ν ≤ f²/N²min."""
    f: Coef
    N2min: float
    is_on: bool = True


def ν_eddy(eddy_param: EddyParameterization, f, αbz, smoothing=10, ν_min=1):
    """inputs.jl:130-137 on arrays of quadrature-point values (``f`` evaluated there too)."""
    import numpy as np
    ν = f * (f / np.sqrt(eddy_param.N2min ** 2 + αbz * αbz))
    return np.logaddexp(smoothing * ν_min, smoothing * ν) / smoothing


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
        pass
