"""Configuration structs: ``Parameters``, surface boundary conditions, ``Forcings``.

Mirrors reference ``src/inputs.jl:3-15`` (Parameters), ``:33-59`` (surface BCs) and ``:141-189``
(Forcings).  Functions of space take an ``(n, 3)`` array of points and return ``(n,)`` values
(the vectorised form of the reference's ``x -> ...`` closures); plain numbers are accepted too.
The convection / eddy parameterisations (``inputs.jl:63-137``) are outside the hot-path scope
 and only carried as switched-off markers.
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
        if self.conv_param.is_on or self.eddy_param.is_on:
            raise NotImplementedError(
                "convection / eddy parameterisations are not on the B200 hot path yet "
                "(SURVEY.md §8 f-2)")
