"""Oracle pinned against /root/reference/test/timesteppers.jl, test/spectral.jl and the
analytic Taylor-Green solution of examples/TaylorGreenVortex2D.jl. CPU only."""
import numpy as np
import pytest

import oracle as O
from oracle import bc, initializers as ini, operators as ops, pressure, timestep as ts

from oracle_helpers import periodic_setup, psolver_cg, rel_l2


def test_tgv2d_analytic_decay():
    # examples/TaylorGreenVortex2D.jl:59-72 (n reduced): u = -sin x cos y e^{-2 nu t}
    n, nu = 32, 5e-2
    setup = periodic_setup(n, 2, L=2 * np.pi)
    psolver = pressure.default_psolver(setup)

    def sol(t):
        return lambda dim, x, y: (
            -np.sin(x) * np.cos(y) if dim == 0 else np.cos(x) * np.sin(y)
        ) * np.exp(-2 * nu * t)

    u0 = ini.velocityfield(setup, sol(0.0), psolver=psolver)
    for method, tol in ((ts.RK44(), 2e-3), (ts.LMWray3(), 2e-3)):
        out = ts.solve_unsteady(
            setup, (0.0, 1.0), dict(u=u0), ts.make_navierstokes(setup, nu), method=method,
            psolver=psolver, dt=0.02,
        )
        uex = ini.velocityfield(setup, sol(1.0), doproject=False)
        for a in range(2):
            assert rel_l2(out["u"][setup.Iu[a] + (a,)], uex[setup.Iu[a] + (a,)]) < tol
        div = ops.divergence_(O.scalarfield(setup), out["u"], setup)
        assert np.max(np.abs(div[setup.Ip])) < 1e-12


def test_wray3_tableau_equals_lmwray3():
    # RKMethods.Wray3 and LMWray3 are the same scheme; with Boussinesq
    # forcing on 16^2 periodic like test/timesteppers.jl:1-42.
    setup = periodic_setup(16, 2, temp=True)
    psolver = pressure.default_psolver(setup)
    rng = np.random.default_rng(3)
    u = ini.random_field(setup, rng=rng, kpeak=4)
    temp = np.asfortranarray(rng.standard_normal(setup.N))
    bc.apply_bc_temp(temp, 0.0, setup)
    force = ts.make_boussinesq(setup, 1e-3, 1e-3, 1, 1.0, True)
    res = []
    for method in (ts.LMWray3(), ts.Wray3()):
        state = dict(u=u.copy(order="F"), temp=temp.copy(order="F"))
        cache = ts.get_cache(method, state, setup)
        st = ts.create_stepper(method, setup, psolver, state, 0.0)
        st = ts.timestep_(method, force, st, 0.1, cache)
        res.append(st["state"])
    assert rel_l2(res[0]["u"], res[1]["u"]) < 1e-13
    assert rel_l2(res[0]["temp"], res[1]["temp"]) < 1e-13


def test_random_field_divergence_free_and_energy():
    # test/spectral.jl:29-69
    for D, n in ((2, 32), (3, 16)):
        setup = periodic_setup(n, D)
        u = ini.random_field(setup, rng=np.random.default_rng(123), kpeak=4, totalenergy=1.0)
        div = ops.divergence_(O.scalarfield(setup), u, setup)
        assert np.max(np.abs(div[setup.Ip])) < 1e-8
        e = sum(np.mean(u[setup.Ip + (a,)] ** 2) for a in range(D)) / 2
        assert np.isclose(e, 1.0, rtol=1e-10)


def test_stage_times_pulsating_plug_flow():
    # test/timesteppers.jl:44-91 (grid/time reduced; default psolver there is
    # psolver_direct -> test-only CG stand-in)
    def uin(dim, x, y, t):
        return (1 + 0.5 * np.sin(np.pi * 2 * t)) + 0 * x + 0 * y if dim == 0 else 0 * x + 0 * y

    bcs = dict(u=((O.DirichletBC(uin), O.PressureBC()), (O.SymmetricBC(), O.SymmetricBC())))
    setup = O.Setup((np.linspace(0, 4, 17), np.linspace(0, 1, 9)), bcs)
    psolver = psolver_cg(setup)
    u0 = O.vectorfield(setup)
    u0[..., 0] = 1.0
    bc.apply_bc_u(u0, 0.0, setup)
    dt, tend = 0.005, 0.1
    for method in (ts.LMWray3(), ts.RK44()):
        out = ts.solve_unsteady(
            setup, (0.0, tend), dict(u=u0), ts.make_navierstokes(setup, 0.05), method=method,
            psolver=psolver, dt=dt,
        )
        uerr = np.max(np.abs(out["u"][setup.Iu[0] + (0,)] - uin(0, 0.0, 0.0, tend)))
        div = ops.divergence_(O.scalarfield(setup), out["u"], setup)
        assert uerr < 1e-10
        assert np.max(np.abs(div[setup.Ip])) < 1e-10


def test_float32_rk44_coefficient_promotion():
    # default RK44() keeps Float64 coefficients even for Float32 fields
    setup = periodic_setup(16, 2, np.float32)
    psolver = pressure.default_psolver(setup)
    u = ini.random_field(setup, rng=np.random.default_rng(0), kpeak=4)
    assert u.dtype == np.float32
    outs = []
    for method in (ts.RK44(), ts.RK44(np.float32)):
        state = dict(u=u.copy(order="F"))
        cache = ts.get_cache(method, state, setup)
        st = ts.create_stepper(method, setup, psolver, state, 0.0)
        st = ts.timestep_(method, ts.make_navierstokes(setup, 1e-3), st, 0.01, cache)
        assert st["state"]["u"].dtype == np.float32
        outs.append(st["state"]["u"])
    assert rel_l2(outs[0], outs[1]) < 1e-6


def test_cfl():
    setup = periodic_setup(16, 3)
    u = ini.random_field(setup, rng=np.random.default_rng(0), kpeak=3)
    dt = ops.propose_timestep(dict(u=u), setup, 1e-3)
    h = 1 / 16
    assert dt <= h**2 / 1e-3 / 6 + 1e-15
    assert np.isclose(dt, min(h / np.max(np.abs(u[setup.inside])), h**2 / 1e-3 / 6))


@pytest.mark.parametrize("name,order", [("FE11", 1), ("Heun33", 3), ("Wray3", 3), ("RK33C2", 3), ("RK44", 4), ("LMWray3", 3)])
def test_observed_temporal_order_of_the_steppers(name, order):
    """The projected Runge-Kutta steppers reach their classical order on a smooth periodic flow (time-independent
    boundary data; Sanderse & Koren 2012): error against a fine-step solution, halving the step three times.  A random
    solenoidal field is used -- for a single Taylor-Green mode the convective term is a gradient and only the linear
    decay would be integrated."""
    import json
    import os
    from fractions import Fraction

    setup = periodic_setup(12, 2, L=2 * np.pi)
    psolver = pressure.default_psolver(setup)
    u0 = ini.random_field(setup, rng=np.random.default_rng(1), kpeak=2, totalenergy=0.5)
    force = ts.make_navierstokes(setup, 0.05)
    if name == "LMWray3":
        method = ts.LMWray3()
    elif hasattr(ts, name):
        method = getattr(ts, name)()
    else:  # from the reference's tableaus (tests/golden/rk_tableaus.json)
        gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rk_tableaus.json")))[name]
        fr = lambda v: float(Fraction(v))
        method = ts.runge_kutta_method([[fr(v) for v in r] for r in gold["A"]], [fr(v) for v in gold["b"]],
                                       [fr(v) for v in gold["c"]], fr(gold["r"]))

    def run(nstep):
        out = ts.solve_unsteady(setup, (0.0, 0.8), dict(u=u0.copy(order="F")), force, method=method, psolver=psolver,
                                dt=0.8 / nstep)
        return out["u"]

    ref = run(256 if order > 1 else 4096)
    errs = [rel_l2(run(n), ref) for n in (4, 8, 16)]
    rates = [np.log2(errs[i] / errs[i + 1]) for i in range(2)]
    assert errs[-1] < errs[0] and min(rates) > order - 0.2, (name, errs, rates)
    assert max(rates) < order + 0.3, (name, errs, rates)
