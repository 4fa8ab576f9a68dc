"""The reference's test and example set-ups as data (BASELINE.json ``configs``).

Each function returns a ``Workload``: parameters, forcings, mesh, spaces arguments, timestepper
and initial buoyancy, restating

* ``bowl_mixing``       — reference ``test/bowl_mixing_tests.jl:15-77`` (config 1)
* ``bowl_wind``         — ``test/bowl_wind_tests.jl:14-64``
* ``bowl_dirichlet``    — ``test/bowl_dirichlet_tests.jl:14-64``
* ``bowl_surface_flux`` — ``test/bowl_surface_flux_tests.jl:14-62``
* ``bowl_example``      — ``examples/bowl_mixing.jl:35-52,83,159,171`` (config 2: h=0.08, μϱ=1,
  BDF2 Δt=1e-3, b(0) = 0.1 exp(−(z+H)/(0.1α)))
* ``channel_basin_box`` — BASELINE config 4 on the declared substitute mesh (structured box of the
  ``meshes/channel_basin_flat.jl`` extent): wind stress of ``test/bowl_wind_tests.jl:27``, surface
  buoyancy flux of ``test/bowl_surface_flux_tests.jl:29``, and the production time stepping of
  ``scratch/run.jl:119-121,163``: ``BDF1(adaptive=true, CFL_factor=0.8)`` with the convection and
  eddy parameterisations switched on
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Any, Callable

import numpy as np

from .dofs import FEData
from .inputs import Forcings, Parameters, SurfaceDirichletBC, SurfaceFluxBC
from .meshes import Mesh
from .spaces import Spaces
from .timesteppers import BDF1, BDF2

MESH_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "meshes")


def mesh_path(dim: int, h: float) -> str:
    return os.path.join(MESH_DIR, f"bowl{dim}D_h{h:.2f}.npz")


@dataclass
class Workload:
    name: str
    params: Parameters
    forcings: Forcings
    mesh: Any                      # path or RawMesh
    spaces_kwargs: dict
    timestepper_kwargs: dict       # for BDF2
    b0: Callable | float | None    # initial buoyancy
    invert_first: bool = False     # examples/bowl_mixing.jl:194 inverts before run!
    tet_rule: str | None = None
    _fe: FEData | None = field(default=None, repr=False)

    def fe_data(self) -> FEData:
        if self._fe is None:
            mesh = Mesh(self.mesh, tet_rule=self.tet_rule)
            self._fe = FEData(mesh, Spaces(mesh, **self.spaces_kwargs))
        return self._fe

    def timestepper(self):
        if self.timestepper_kwargs.get("adaptive"):
            return BDF1(**self.timestepper_kwargs)
        return BDF2(**self.timestepper_kwargs)


_U_DIRI = dict(u_diri_tags=["bottom", "coastline", "surface"],
               u_diri_vals=[(0, 0, 0)] * 3,
               u_diri_masks=[(True, True, True), (True, True, True), (False, False, True)])


def _H(α):
    return lambda x: α * (1 - x[:, 0] ** 2 - x[:, 1] ** 2)


def _kappa_bottom(α):
    H = _H(α)
    return lambda x: 1e-2 + np.exp(-(x[:, 2] + H(x)) / (0.1 * α))


def bowl_mixing(dim: int = 3, h: float = 0.1, mesh=None) -> Workload:
    ε, α, μϱ = 2e-1, 0.5, 1e1
    params = Parameters(ε=ε, α=α, μϱ=μϱ, N2=1 / α, f=lambda x: 1 + 0.5 * x[:, 1], H=_H(α))
    κ = _kappa_bottom(α)
    forcings = Forcings(1, κ, κ, 0.0, 0.0, SurfaceDirichletBC(0.0))
    Δt = 1e-4 * μϱ / (α * ε) ** 2
    return Workload(f"bowl_mixing_{dim}D", params, forcings, mesh or mesh_path(dim, h),
                    dict(_U_DIRI, b_diri_tags=["coastline", "surface"], b_diri_vals=[0.0, 0.0]),
                    dict(t_start=0.0, t_stop=50 * Δt, Δt=Δt), None)


def _wind_like(name, κ, τx, b_surface, b0):
    ε, α = np.sqrt(1e-1), 0.5
    params = Parameters(ε=ε, α=α, μϱ=1, N2=0, f=lambda x: 0 + 0.5 * x[:, 1], H=_H(α))
    forcings = Forcings(1, κ, κ, τx, 0.0, SurfaceDirichletBC(b_surface))
    return Workload(name, params, forcings, mesh_path(3, 0.1),
                    dict(_U_DIRI, b_diri_tags=["coastline", "surface"],
                         b_diri_vals=[b_surface, b_surface]),
                    dict(t_start=0.0, t_stop=50 * 1e-1, Δt=1e-1), b0)


def bowl_wind() -> Workload:
    α = 0.5
    return _wind_like("bowl_wind", _kappa_bottom(α), lambda x: -1e-1 * np.cos(np.pi * x[:, 1] / 2),
                      0.0, lambda x: x[:, 2] / α)


def bowl_dirichlet() -> Workload:
    b_surface = lambda x: x[:, 1]                                   # noqa: E731
    return _wind_like("bowl_dirichlet", 1.0, 0.0, b_surface, b_surface)


def bowl_surface_flux() -> Workload:
    ε, α = np.sqrt(1e-1), 0.5
    params = Parameters(ε=ε, α=α, μϱ=1, N2=0, f=lambda x: np.ones(len(x)), H=_H(α))
    forcings = Forcings(1, 1e-2, 1e-2, 0.0, 0.0,
                        SurfaceFluxBC(lambda x: 1e-3 * np.sin(np.pi * x[:, 0])))
    return Workload("bowl_surface_flux", params, forcings, mesh_path(3, 0.1), dict(_U_DIRI),
                    dict(t_start=0.0, t_stop=50 * 1e-1, Δt=1e-1), lambda x: x[:, 2] / α)


def refined_bowl(levels: int = 1, h0: float = 0.08, α: float = 0.5):
    """bowl3D mesh refined `levels` times from the shipped h0 mesh (h0/2, h0/4, ...), boundary
    nodes projected onto the bowl: stands in for the gmsh-generated meshes of BASELINE configs
    3 (h = 0.04) and 5 (h = 0.02)."""
    from .gridap_lite import RawMesh, bowl_projection, refine
    raw = RawMesh.load_npz(mesh_path(3, h0))
    for _ in range(levels):
        raw = refine(raw, project=bowl_projection(α))
    return raw


def bowl_example(h: float = 0.08, mesh=None, n_steps: int | None = None) -> Workload:
    ε, α, μϱ = 2e-1, 0.5, 1.0
    H = _H(α)
    params = Parameters(ε=ε, α=α, μϱ=μϱ, N2=1 / α, f=lambda x: 1.0 + 0.5 * x[:, 1], H=H)
    κ = _kappa_bottom(α)
    forcings = Forcings(1, κ, κ, 0.0, 0.0, SurfaceDirichletBC(0.0))
    Δt = 1e-3
    t_stop = 0.1 * μϱ / ε ** 2 if n_steps is None else n_steps * Δt
    return Workload(f"bowl_example_h{h:g}", params, forcings, mesh or mesh_path(3, h),
                    dict(_U_DIRI, b_diri_tags=["coastline", "surface"], b_diri_vals=[0.0, 0.0]),
                    dict(t_start=0.0, t_stop=t_stop, Δt=Δt),
                    lambda x: 0.1 * np.exp(-(x[:, 2] + H(x)) / (0.1 * α)), invert_first=True)


def channel_basin_box(n=(6, 12, 4), α: float = 0.125, periodic: bool = False) -> Workload:
    """Config 4 substitute: flat channel-basin box x∈[0,1], y∈[−1,1], z∈[−α,0] (n cells per
    direction), wind + surface flux forcing, adaptive BDF1, convection + eddy parameterisations.
    ``periodic``: the channel part y <= −1/2 is periodic in x, as in ``meshes/channel_basin_flat.jl:5-10,114-121``
    (the basin part keeps its walls)."""
    from .gridap_lite import box_mesh
    from .inputs import ConvectionParameterization, EddyParameterization
    ε, μϱ = np.sqrt(1e-1), 1.0
    f = lambda x: 1.0 + 0.5 * x[:, 1]                                           # noqa: E731
    H = lambda x: np.full(len(x), α)                                            # noqa: E731
    params = Parameters(ε=ε, α=α, μϱ=μϱ, N2=1.0, f=f, H=H)
    κ = lambda x: 1e-2 + np.exp(-(x[:, 2] + α) / (0.1 * α))                     # noqa: E731
    forcings = Forcings(1, κ, κ, lambda x: -1e-1 * np.cos(np.pi * x[:, 1] / 2), 0.0,
                        SurfaceFluxBC(lambda x: 1e-3 * np.sin(np.pi * x[:, 0])),
                        conv_param=ConvectionParameterization(κᶜ=1.0, N2min=1e-3),
                        eddy_param=EddyParameterization(f=f, N2min=np.sqrt(1e-3)))
    mesh = box_mesh(*n, z=(-α, 0.0), periodic_x_below=-0.5 if periodic else None)
    return Workload("channel_basin_box" + ("_periodic" if periodic else ""), params, forcings, mesh, dict(_U_DIRI),
                    # CFL_factor: the reference's production value is 0.8 on its h = 1e-2 mesh; on this
                    # coarse stand-in the u_min = 0.01 floor alone would give Δt ≈ 19 at rest, so 0.05
                    dict(t_start=0.0, t_stop=float("inf"), Δt=1e-2, adaptive=True, CFL_factor=0.05),
                    lambda x: 0.1 * x[:, 2] / α)


def with_b_order(w: Workload, b_order: int) -> Workload:
    """The same set-up with buoyancy of order ``b_order`` (``Spaces(...; b_order=1)`` is what the
    production runs use, ``scratch/run.jl:152``)."""
    w.spaces_kwargs = dict(w.spaces_kwargs, b_order=b_order)
    w.name = f"{w.name}_bP{b_order}"
    w._fe = None
    return w


def host_operands(w: Workload) -> dict:
    """Every host-side operand of the solve path, in solver (permuted) order — what a Julia host
    would hand over the C ABI.  Also the input of the CPU oracle."""
    from .element_tables import element_tables
    from .evolution import permuted_evolution_system
    from .inversion import permuted_inversion_system
    fe = w.fe_data()
    A, B, b0, pscale = permuted_inversion_system(fe, w.params, w.forcings)
    ops = permuted_evolution_system(fe, w.params, w.forcings)
    ops.update(A=A, B=B, b0=b0, pscale=pscale, tables=element_tables(fe),
               nu=fe.dofs.nu, np=fe.dofs.np, nb=fe.dofs.nb)
    if w.b0 is None:
        ops["b_init"] = np.zeros(fe.dofs.nb)
    else:
        ops["b_init"] = fe.spaces.B.interpolate(w.b0)[0][fe.dofs.p_b]
    return ops
