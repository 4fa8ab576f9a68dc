"""Host-side mirror of the reference API (no GPU needed): timesteppers, parameterisation formulas,
forcings, block-preconditioner operands."""
import numpy as np
import pytest

from conftest import workload
from nupgcm_b200.inputs import (ConvectionParameterization, EddyParameterization, Forcings,
                                SurfaceDirichletBC, κᵥ_convection, ν_eddy)
from nupgcm_b200.timesteppers import BDF1, BDF2, evolution_parameter, update_t_, update_Δt_


def test_timesteppers_mirror_the_reference():
    bdf1 = BDF1(t_start=0.0, t_stop=1.0, Δt=0.1, adaptive=True, CFL_factor=0.5)
    bdf2 = BDF2(t_start=0.0, t_stop=1.0, Δt=0.1)
    assert bdf1.adaptive and bdf1.scheme == 1 and not bdf2.adaptive and bdf2.scheme == 2
    update_Δt_(bdf2)                                          # no-op for BDF2 (timesteppers.jl:120-122)
    assert bdf2.Δt == 0.1

    class FakeMesh:
        def cfl_dt(self, u, cfl, u_min):
            return cfl * 2.0 / u_min

    update_Δt_(bdf1, FakeMesh(), None)
    assert bdf1.Δt == pytest.approx(0.5 * 2.0 / 0.01)
    # declared deviation from timesteppers.jl:108-119 (see update_Δt_): a non-adaptive BDF1 keeps its Δt
    fixed = BDF1(t_start=0.0, t_stop=1.0, Δt=0.1)
    update_Δt_(fixed, FakeMesh(), None)
    assert fixed.Δt == 0.1
    update_t_(bdf1)
    assert bdf1.t == pytest.approx(100.0)
    # t_stop = 50 Δt takes 51 steps with `while t < t_stop; t += Δt`
    ts = BDF2(t_start=0.0, t_stop=50 * 0.1, Δt=0.1)
    n = 0
    while ts.t < ts.t_stop:
        update_t_(ts)
        n += 1
    assert n == 51

    class P:
        α, ε, μϱ = 0.5, 0.2, 10.0
    assert evolution_parameter(P, bdf2) == pytest.approx(2 / 3 * 0.1 * 0.25 * 0.04 / 10.0)
    bdf1.Δt = 0.1
    assert evolution_parameter(P, bdf1) == pytest.approx(0.1 * 0.25 * 0.04 / 10.0)


def test_parameterisation_formulas():
    conv = ConvectionParameterization(κᶜ=2.0, N2min=1e-3)
    # stable stratification -> no extra mixing; unstable -> κᶜ; neutral -> κᶜ/2 (inputs.jl:87-91)
    assert κᵥ_convection(conv, 0.1, np.array([1.0]))[0] == pytest.approx(0.1, abs=1e-12)
    assert κᵥ_convection(conv, 0.1, np.array([-1.0]))[0] == pytest.approx(2.1, abs=1e-12)
    assert κᵥ_convection(conv, 0.1, np.array([0.0]))[0] == pytest.approx(1.1)
    eddy = EddyParameterization(f=1.0, N2min=0.1)
    f = np.array([1.0, 2.0])
    # weak stratification: capped at f²/N²min; strong stratification: floored at ν_min (inputs.jl:130-137)
    assert np.allclose(ν_eddy(eddy, f, np.zeros(2)), f * f / 0.1, rtol=1e-6)
    assert np.allclose(ν_eddy(eddy, f, np.full(2, 1e6)), 1.0, rtol=1e-5)       # smooth maximum: 1 + log1p(e^-10)/10
    # same formula as the oracle's restatement (zero buoyancy, N² = 0.3)
    from oracle.element_rhs import nu_friction  # noqa: F401  (import check only: formulas are compared on the GPU tests)
    fo = Forcings(1, 1e-2, 1e-2, 0.0, 0.0, SurfaceDirichletBC(0.0), conv_param=conv, eddy_param=eddy)
    assert fo.conv_param.is_on and fo.eddy_param.is_on
    assert not Forcings(1, 1e-2, 1e-2, 0.0, 0.0, SurfaceDirichletBC(0.0)).conv_param.is_on


def test_block_preconditioner_operands():
    """P = friction block with ν = 1 (SPD, nu x nu, explicit zeros dropped), T = pressure mass /
    (α²ε²) (SPD, np x np) — preconditioners.jl:62-93."""
    from nupgcm_b200.preconditioners import block_operands
    w, ops = workload("bowl_mixing", dim=2)
    fe = w.fe_data()
    F, T = block_operands(w.params, fe)
    assert F.shape == (fe.dofs.nu, fe.dofs.nu) and T.shape == (fe.dofs.np, fe.dofs.np)
    assert abs(F - F.T).max() < 1e-15 and abs(T - T.T).max() < 1e-15
    assert np.count_nonzero(F.data) == F.nnz
    assert np.linalg.eigvalsh(T.toarray()).min() > 0 and np.linalg.eigvalsh(F.toarray()).min() > 0
    # the friction block of the assembled inversion matrix (constant ν = 1) is exactly P
    A = ops["A"].tocsr()[:fe.dofs.nu, :fe.dofs.nu]
    sym = ((A + A.T) / 2).tocsr()                              # Coriolis is the skew part
    d = (sym - F).tocsr()
    assert abs(d.data).max() < 1e-12 * abs(F.data).max()
