"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Tolerances (BASELINE.json north_star): relative L2 per step <= 1e-12 in Float64 and
<= 1e-5 in Float32; ghost fills are bit-exact; post-projection divergence at machine
precision."""
import numpy as np
import pytest

import oracle as O
from oracle import bc as obc, operators as oops, pressure as opr, timestep as ots, initializers as oini

from gpu_helpers import make_setups, random_scal, random_vec, rel_l2, smooth_vec, uniform

pytestmark = pytest.mark.gpu

TOL = {np.float64: 1e-12, np.float32: 1e-5}
OPTOL = {np.float64: 5e-13, np.float32: 3e-5}  # single operator with cancellation (k = O(n) * eps)


def ins():
    import insb200

    return insb200


GRIDS3 = [
    ("periodic", (uniform(16), uniform(12, 2.0), uniform(20, 0.5)), ("P", "P", "P")),
    ("walls_stretched", (O.tanh_grid(0, 1, 12, 1.2), O.tanh_grid(0, 1, 10, 1.1), O.cosine_grid(0, 1, 14)), ("D", "D", "D")),
    ("channel", (uniform(16, 4.0), O.tanh_grid(0, 2, 12, 1.4), uniform(8, 2.0)), ("P", "D", "P")),
    ("mixed", (uniform(10), O.cosine_grid(0, 1, 8), uniform(12)), (("D", "O"), "S", "P")),
    ("lid", (uniform(8), uniform(10), uniform(6)), ("P", (("D", (0.0, 0.0, 0.0)), ("D", (1.0, 0.0, 0.5))), "S")),
]
GRIDS2 = [
    ("periodic2", (uniform(16), uniform(24, 2.0)), ("P", "P")),
    ("walls2", (O.tanh_grid(0, 1, 16), O.tanh_grid(0, 1, 12, 1.3)), ("D", "D")),
    ("mixed2", (uniform(12, 4.0), uniform(8)), (("D", "O"), "S")),
]


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("name,x,bcs", GRIDS3 + GRIDS2)
def test_bc_fills_bit_exact(name, x, bcs, T):
    I = ins()
    tb = tuple("P" if b == "P" else ("S" if b == "S" else (("D", 1.0), ("D", 0.0)) if b == "D" else "S") for b in bcs)
    so, sg = make_setups(x, bcs, temp_bcs=tb, T=T)
    u = random_vec(so, 1)
    p = random_scal(so, 2)
    tmp = random_scal(so, 3)
    ug = I.apply_bc_u_(I.to_device(sg, u), 0.0, sg)
    pg = I.apply_bc_p_(I.to_device(sg, p), 0.0, sg)
    tg = I.apply_bc_temp_(I.to_device(sg, tmp), 0.0, sg)
    assert np.array_equal(I.to_host(ug), obc.apply_bc_u(u.copy(order="F"), 0.0, so))
    assert np.array_equal(I.to_host(pg), obc.apply_bc_p(p.copy(order="F"), 0.0, so))
    assert np.array_equal(I.to_host(tg), obc.apply_bc_temp(tmp.copy(order="F"), 0.0, so))
    # dudt = true: constant Dirichlet values become zero (boundary_conditions.jl:347)
    ug2 = I.apply_bc_u_(I.to_device(sg, u), 0.0, sg, dudt=True)
    assert np.array_equal(I.to_host(ug2), obc.apply_bc_u(u.copy(order="F"), 0.0, so, dudt=True))


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("name,x,bcs", GRIDS3 + GRIDS2)
def test_operators_match_oracle(name, x, bcs, T):
    I = ins()
    so, sg = make_setups(x, bcs, temp_bcs=tuple("P" if b == "P" else "S" for b in bcs), T=T)
    D = so.D
    u = smooth_vec(so, 4)
    p = obc.apply_bc_p(random_scal(so, 5), 0.0, so)
    temp = obc.apply_bc_temp(random_scal(so, 6), 0.0, so)
    ud, pd, td = I.to_device(sg, u), I.to_device(sg, p), I.to_device(sg, temp)
    tol = OPTOL[T]
    inside = so.inside
    # divergence / scalewithvolume
    div_o = oops.divergence_(O.scalarfield(so), u, so)
    div_g = I.to_host(I.divergence(ud, sg))
    assert rel_l2(div_g[inside], div_o[inside]) < tol
    sc_o = oops.scalewithvolume_(p.copy(order="F"), so)
    assert rel_l2(I.to_host(I.scalewithvolume_(pd.clone(), sg)), sc_o) < 1e-6 if T is np.float32 else 1e-15
    # pressure gradient / applypressure
    G_o = oops.pressuregradient_(O.vectorfield(so), p, so)
    assert rel_l2(I.to_host(I.pressuregradient(pd, sg)), G_o) < tol
    up_o = oops.applypressure_(u.copy(order="F"), p, so)
    assert rel_l2(I.to_host(I.applypressure_(ud.clone(), pd, sg)), up_o) < tol
    # convection / diffusion add into f
    f0 = random_vec(so, 7)
    c_o = oops.convection_(f0.copy(order="F"), u, so)
    c_g = I.to_host(I.convection_(I.to_device(sg, f0), ud, sg))
    assert rel_l2(c_g, c_o) < tol
    d_o = oops.diffusion_(f0.copy(order="F"), u, so, 1e-2)
    d_g = I.to_host(I.diffusion_(I.to_device(sg, f0), ud, sg, 1e-2))
    assert rel_l2(d_g, d_o) < tol
    # navierstokes! and boussinesq!
    k_o = {"u": O.vectorfield(so), "temp": O.scalarfield(so)}
    oops.navierstokes_(k_o, {"u": u}, 0.0, so, 1e-2)
    k_g = {"u": I.to_device(sg, random_vec(so, 8)), "temp": I.to_device(sg, random_scal(so, 9))}
    I.navierstokes_(k_g, {"u": ud}, 0.0, setup=sg, viscosity=1e-2)
    assert rel_l2(I.to_host(k_g["u"]), k_o["u"]) < tol
    for diss in (False, True):
        oops.boussinesq_(k_o, {"u": u, "temp": temp}, 0.0, so, 1e-2, 2e-2, D - 1, 0.7, diss)
        I.boussinesq_(k_g, {"u": ud, "temp": td}, 0.0, setup=sg, viscosity=1e-2, conductivity=2e-2,
                      gdir=D - 1, gravity=0.7, dodissipation=diss)
        assert rel_l2(I.to_host(k_g["u"]), k_o["u"]) < tol
        assert rel_l2(I.to_host(k_g["temp"]), k_o["temp"]) < tol
    # CFL and max|u|
    dt_o = oops.propose_timestep({"u": u}, so, 1e-2, 2e-2)
    dt_g = I.propose_timestep(I.boussinesq_, {"u": ud}, sg,
                              dict(viscosity=1e-2, conductivity=2e-2, gravity=0, gdir=0, dodissipation=False))
    assert np.isclose(float(dt_g), float(dt_o), rtol=1e-6 if T is np.float32 else 1e-14)
    assert np.isclose(I.maxabs(ud, sg), np.max(np.abs(u)))


PSOLVER_CASES = [
    # name, x, bcs, kind
    ("spectral2", (uniform(32, 2 * np.pi), uniform(32, 2 * np.pi)), ("P", "P"), "spectral"),
    ("spectral3", (uniform(16), uniform(12, 2.0), uniform(20, 0.5)), ("P", "P", "P"), "spectral"),
    ("spectral3b", (uniform(64), uniform(32), uniform(16)), ("P", "P", "P"), "spectral"),
    # every length of the register-FFT fast path (32..2048), in every position
    ("spectral_fast_128_32_256", (uniform(128), uniform(32), uniform(256)), ("P", "P", "P"), "spectral"),
    ("spectral_fast_256_128_32", (uniform(256), uniform(128), uniform(32)), ("P", "P", "P"), "spectral"),
    ("spectral_fast_512_32_64", (uniform(512), uniform(32), uniform(64)), ("P", "P", "P"), "spectral"),
    ("spectral_fast_32_1024_32", (uniform(32), uniform(1024), uniform(32)), ("P", "P", "P"), "spectral"),
    ("spectral_fast_64_32_512", (uniform(64), uniform(32), uniform(512)), ("P", "P", "P"), "spectral"),
    ("spectral_fast_32_32_2048", (uniform(32), uniform(32), uniform(2048)), ("P", "P", "P"), "spectral"),
    ("spectral_fast_2048_64", (uniform(2048, 2 * np.pi), uniform(64, 2 * np.pi)), ("P", "P"), "spectral"),
    ("spectral_fast_1024_2048", (uniform(1024, 2 * np.pi), uniform(2048, 2 * np.pi)), ("P", "P"), "spectral"),
    ("transform_pp", (uniform(32, 2 * np.pi), uniform(32, 2 * np.pi)), ("P", "P"), "transform"),
    ("transform_pw", (uniform(16, 2.0), uniform(12)), ("P", "D"), "transform"),
    ("transform_wp", (uniform(10), uniform(16, 2.0)), ("D", "P"), "transform"),
    ("transform_ww", (uniform(20), uniform(26)), ("D", "D"), "transform"),
    ("transform_pww", (uniform(16, np.pi), uniform(12), uniform(10)), ("P", "D", "D"), "transform"),
    ("transform_wpw", (uniform(10), uniform(16), uniform(28)), ("D", "P", "D"), "transform"),
    ("transform_www", (uniform(10), uniform(16), uniform(28)), ("D", "D", "D"), "transform"),
    ("transform_ppp", (uniform(8), uniform(16), uniform(4)), ("P", "P", "P"), "transform"),
    # register-FFT fast paths of psolver_transform (x periodic, power-of-two extents)
    ("transform_fast_pww", (uniform(64, np.pi), uniform(32), uniform(128)), ("P", "D", "D"), "transform"),
    ("transform_fast_ppw", (uniform(32), uniform(64), uniform(32)), ("P", "P", "D"), "transform"),
    ("transform_fast_pwp", (uniform(128), uniform(256), uniform(32)), ("P", "D", "P"), "transform"),
    ("transform_fast_pw", (uniform(64, 2.0), uniform(512)), ("P", "D"), "transform"),
    # ... and of psolver_tridiagonal (walls in y or z, marching Thomas pass)
    ("tridiag_fast_y", (uniform(64, 4.0), O.tanh_grid(0, 2, 40, 1.4), uniform(128, 2.0)), ("P", "D", "P"), "tridiagonal"),
    ("tridiag_fast_z", (uniform(128, 4.0), uniform(32, 2.0), O.cosine_grid(0, 1, 23)), ("P", "P", "D"), "tridiagonal"),
]
for _D in (2, 3):
    for _dir in range(_D):
        _x = tuple(O.cosine_grid(0.0, 1.0, 16) if b == _dir else uniform(32, 2 * np.pi) for b in range(_D))
        _b = tuple("D" if b == _dir else "P" for b in range(_D))
        PSOLVER_CASES.append((f"tridiag_D{_D}_dir{_dir}", _x, _b, "tridiagonal"))


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("name,x,bcs,kind", PSOLVER_CASES)
def test_psolvers_match_oracle(name, x, bcs, kind, T):
    I = ins()
    so, sg = make_setups(x, bcs, T=T)
    mk_o = {"spectral": opr.psolver_spectral, "transform": opr.psolver_transform, "tridiagonal": opr.psolver_tridiagonal}[kind]
    mk_g = {"spectral": I.psolver_spectral, "transform": I.psolver_transform, "tridiagonal": I.psolver_tridiagonal}[kind]
    # right-hand side in the range of the Laplacian: volume-scaled divergence of a field with BCs
    u = smooth_vec(so, 11)
    f = oops.scalewithvolume_(oops.divergence_(O.scalarfield(so), u, so), so)
    ps_o, ps_g = mk_o(so), mk_g(sg)
    assert ps_g.kind == kind
    p_o = ps_o(f.copy(order="F"))
    p_g = I.to_host(ps_g(I.to_device(sg, f)))
    Ip = so.Ip
    tol = 1e-12 if T is np.float64 else 2e-5
    assert rel_l2(p_g[Ip], p_o[Ip]) < tol, name
    # untouched outside Ip
    mask = np.ones(so.N, bool)
    mask[Ip] = False
    assert np.array_equal(p_g[mask], f[mask])
    # manufactured solution: L p = lap  =>  solver recovers p up to a constant (test/psolvers.jl)
    pe = O.scalarfield(so)
    pe[Ip] = np.random.default_rng(3).standard_normal(so.Np).astype(T)
    obc.apply_bc_p(pe, 0.0, so)
    lap = oops.laplacian_(O.scalarfield(so), pe, so)
    got = I.to_host(ps_g(I.to_device(sg, lap)))[Ip]
    a = got - got.mean()
    b = pe[Ip] - pe[Ip].mean()
    assert rel_l2(a, b) < (1e-9 if T is np.float64 else 5e-3)


def test_default_psolver_choice():
    I = ins()
    _, sg = make_setups((uniform(16), uniform(16)), ("P", "P"))
    assert I.default_psolver(sg).kind == "spectral"
    _, sg = make_setups((uniform(16), O.cosine_grid(0, 1, 8), uniform(8)), ("P", "D", "P"))
    assert I.default_psolver(sg).kind == "tridiagonal"
    _, sg = make_setups((uniform(16), uniform(8)), ("D", "D"))
    assert I.default_psolver(sg).kind == "cg"  # the reference would build psolver_direct here
    with pytest.raises(I.Insb200Error):
        I.psolver_spectral(sg)
    _, sg = make_setups((uniform(16), uniform(16)), ("P", "P"))
    with pytest.raises(I.Insb200Error):
        I.psolver_tridiagonal(sg)


STEP_CASES = [
    ("periodic3", (uniform(16), uniform(12, 2.0), uniform(20, 0.5)), ("P", "P", "P"), "spectral"),
    ("periodic2", (uniform(32), uniform(32)), ("P", "P"), "spectral"),
    ("channel3", (uniform(16, 4.0), O.tanh_grid(0, 2, 12, 1.4), uniform(8, 2.0)), ("P", "D", "P"), "tridiagonal"),
    ("rb3", (uniform(16, np.pi), uniform(12), uniform(10)), ("P", "D", "D"), "transform"),
    ("channel2", (uniform(32, 4.0), O.cosine_grid(0, 1, 16)), ("P", "D"), "tridiagonal"),
]
METHODS = ["RK44", "LMWray3", "Wray3", "FE11", "RK33C2"]


def _methods(mod, name, T=np.float64):
    if name == "LMWray3":
        return mod.LMWray3(T)
    src = mod if hasattr(mod, name) else mod.RKMethods
    return getattr(src, name)(T)


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("mname", METHODS)
@pytest.mark.parametrize("name,x,bcs,kind", STEP_CASES)
def test_timestep_matches_oracle(name, x, bcs, kind, mname, T):
    I = ins()
    tb = tuple("P" if b == "P" else (("D", 1.0), ("D", 0.0)) for b in bcs)
    so, sg = make_setups(x, bcs, temp_bcs=tb, T=T)
    D = so.D
    mk_o = {"spectral": opr.psolver_spectral, "transform": opr.psolver_transform, "tridiagonal": opr.psolver_tridiagonal}[kind]
    mk_g = {"spectral": I.psolver_spectral, "transform": I.psolver_transform, "tridiagonal": I.psolver_tridiagonal}[kind]
    ps_o, ps_g = mk_o(so), mk_g(sg)
    u0 = smooth_vec(so, 21)
    u0 = opr.project(u0, so, ps_o)
    obc.apply_bc_u(u0, 0.0, so)
    temp0 = obc.apply_bc_temp(random_scal(so, 22), 0.0, so)
    dt = 0.01
    for bouss in (False, True):
        params = dict(viscosity=1e-2, conductivity=2e-2, gdir=D - 1, gravity=1.0, dodissipation=True)
        if bouss:
            f_o = ots.make_boussinesq(so, **params)
            f_g, p_g = I.boussinesq_, params
            st_o = {"u": u0.copy(order="F"), "temp": temp0.copy(order="F")}
            st_g = {"u": I.to_device(sg, u0), "temp": I.to_device(sg, temp0)}
        else:
            f_o = ots.make_navierstokes(so, 1e-2)
            f_g, p_g = I.navierstokes_, dict(viscosity=1e-2)
            st_o = {"u": u0.copy(order="F")}
            st_g = {"u": I.to_device(sg, u0)}
        m_o, m_g = _methods(ots, mname), _methods(I, mname)
        cache = ots.get_cache(m_o, st_o, so)
        sp_o = ots.create_stepper(m_o, so, ps_o, st_o, 0.0)
        sp_g = I.create_stepper(m_g, setup=sg, psolver=ps_g, state=st_g, t=0.0)
        for step in range(2):
            sp_o = ots.timestep_(m_o, f_o, sp_o, dt, cache)
            sp_g = I.timestep_(m_g, f_g, sp_g, dt, params=p_g)
            e = rel_l2(I.to_host(sp_g["state"]["u"]), sp_o["state"]["u"])
            assert e < TOL[T] * (step + 1), (name, mname, bouss, step, e)
            if bouss:
                e = rel_l2(I.to_host(sp_g["state"]["temp"]), sp_o["state"]["temp"])
                assert e < TOL[T] * (step + 1), (name, mname, "temp", step, e)
        assert np.isclose(sp_g["t"], sp_o["t"]) and sp_g["n"] == 2
        # post-projection divergence at machine precision
        div = I.to_host(I.divergence(sp_g["state"]["u"], sg))[so.Ip]
        scale = np.max(np.abs(u0)) / min(np.min(d[1:-1]) for d in so.Delta)
        assert np.max(np.abs(div)) < (1e-12 if T is np.float64 else 2e-4) * scale


def test_float32_fields_with_float64_coefficients():
    # default RK44() carries Float64 coefficients: combine in Float64, round once
    I = ins()
    so, sg = make_setups((uniform(16), uniform(16)), ("P", "P"), T=np.float32)
    ps_o, ps_g = opr.psolver_spectral(so), I.psolver_spectral(sg)
    u0 = oini.random_field(so, rng=np.random.default_rng(0), kpeak=4)
    for mT in (np.float64, np.float32):
        m_o, m_g = ots.RK44(mT), I.RKMethods.RK44(mT)
        st_o = {"u": u0.copy(order="F")}
        sp_o = ots.create_stepper(m_o, so, ps_o, st_o, 0.0)
        sp_o = ots.timestep_(m_o, ots.make_navierstokes(so, 1e-3), sp_o, 0.01, ots.get_cache(m_o, st_o, so))
        sp_g = I.create_stepper(m_g, setup=sg, psolver=ps_g, state={"u": I.to_device(sg, u0)}, t=0.0)
        sp_g = I.timestep_(m_g, I.navierstokes_, sp_g, 0.01, params=dict(viscosity=1e-3))
        assert rel_l2(I.to_host(sp_g["state"]["u"]), sp_o["state"]["u"]) < 1e-5


def test_tgv2d_config1_against_oracle_and_analytic():
    # BASELINE config 1: TaylorGreenVortex2D 256^2 periodic Float64, RK44, spectral, 100 steps
    I = ins()
    n, nu = 256, 5e-4
    so, sg = make_setups((uniform(n, 2 * np.pi), uniform(n, 2 * np.pi)), ("P", "P"))

    def sol(t):
        return lambda dim, x, y: (-np.sin(x) * np.cos(y) if dim == 0 else np.cos(x) * np.sin(y)) * np.exp(-2 * nu * t)

    ps_o, ps_g = opr.psolver_spectral(so), I.psolver_spectral(sg)
    u0_o = oini.velocityfield(so, sol(0.0), psolver=ps_o)
    u0_g = I.velocityfield(sg, sol(0.0), psolver=ps_g)
    assert rel_l2(I.to_host(u0_g), u0_o) < 1e-13
    out_o = ots.solve_unsteady(so, (0.0, 2.0), {"u": u0_o}, ots.make_navierstokes(so, nu), method=ots.RK44(),
                               psolver=ps_o, dt=0.02)
    log = []
    out_g, outputs = I.solve_unsteady(setup=sg, tlims=(0.0, 2.0), start={"u": u0_g}, method=I.RKMethods.RK44(),
                                      psolver=ps_g, dt=0.02, params=dict(viscosity=nu),
                                      processors={"log": I.timelogger(sg, nupdate=10, out=log)})
    assert len(log) == 10 and log[-1]["n"] == 100
    ug = I.to_host(out_g["u"])
    assert rel_l2(ug, out_o["u"]) < 1e-11  # 100 steps
    uex = oini.velocityfield(so, sol(2.0), doproject=False)
    assert rel_l2(ug[so.inside], uex[so.inside]) < 1e-3


def test_adaptive_solve_matches_oracle():
    I = ins()
    so, sg = make_setups((uniform(16), uniform(16), uniform(16)), ("P", "P", "P"))
    ps_o, ps_g = opr.psolver_spectral(so), I.psolver_spectral(sg)
    u0 = oini.random_field(so, rng=np.random.default_rng(5), kpeak=3)
    out_o = ots.solve_unsteady(so, (0.0, 0.05), {"u": u0}, ots.make_navierstokes(so, 2e-3), psolver=ps_o, cfl=0.9)
    out_g, _ = I.solve_unsteady(setup=sg, tlims=(0.0, 0.05), start={"u": I.to_device(sg, u0)}, psolver=ps_g,
                                cfl=0.9, params=dict(viscosity=2e-3))
    assert np.isclose(out_g["t"], out_o["t"])
    assert rel_l2(I.to_host(out_g["u"]), out_o["u"]) < 1e-10


def test_right_hand_side_sciml():
    I = ins()
    so, sg = make_setups((uniform(16), O.cosine_grid(0, 1, 12), uniform(8)), ("P", "D", "P"))
    ps_o, ps_g = opr.psolver_tridiagonal(so), I.psolver_tridiagonal(sg)
    u = smooth_vec(so, 30)
    d_o = O.vectorfield(so)
    ots.right_hand_side_(d_o, u, so, ps_o, 1e-2, 0.0)
    d_g = I.vectorfield(sg)
    I.right_hand_side_(d_g, I.to_device(sg, u), (sg, ps_g, 1e-2), 0.0)
    # boundary-value cells hold transient junk in both: compare the DOFs
    got = I.to_host(d_g)
    for a in range(3):
        assert rel_l2(got[so.Iu[a] + (a,)], d_o[so.Iu[a] + (a,)]) < 1e-12


@pytest.mark.parametrize("T", [np.float64, np.float32])
def test_medium_3d_step(T):
    # 64^3 TGV3D-like (BASELINE config 2 shape, reduced): one RK44 step against the oracle
    I = ins()
    n = 64
    so, sg = make_setups((uniform(n), uniform(n), uniform(n)), ("P", "P", "P"), T=T)
    ps_o, ps_g = opr.psolver_spectral(so), I.psolver_spectral(sg)

    def U(dim, x, y, z):
        if dim == 0:
            return np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y) * np.sin(2 * np.pi * z) / 2
        if dim == 1:
            return -np.cos(2 * np.pi * x) * np.sin(2 * np.pi * y) * np.sin(2 * np.pi * z) / 2
        return 0 * x + 0 * y + 0 * z

    u0 = oini.velocityfield(so, U, psolver=ps_o)
    m_o, m_g = ots.RK44(), I.RKMethods.RK44()
    st_o = {"u": u0.copy(order="F")}
    sp_o = ots.create_stepper(m_o, so, ps_o, st_o, 0.0)
    sp_g = I.create_stepper(m_g, setup=sg, psolver=ps_g, state={"u": I.to_device(sg, u0)}, t=0.0)
    cache = ots.get_cache(m_o, st_o, so)
    for _ in range(2):
        sp_o = ots.timestep_(m_o, ots.make_navierstokes(so, 1e-3), sp_o, 2e-3, cache)
        sp_g = I.timestep_(m_g, I.navierstokes_, sp_g, 2e-3, params=dict(viscosity=1e-3))
    assert rel_l2(I.to_host(sp_g["state"]["u"]), sp_o["state"]["u"]) < 2 * TOL[T]
