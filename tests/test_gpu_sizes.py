"""GPU parity at the sizes and on the code paths the benchmarks run ("next round" item 1).

The small grids of test_gpu_parity.py keep the marching stage kernel on one x tile and one z chunk;
here every case forces several tiles / chunks, the unaligned variant, and the BASELINE configs at
(or near) size: one Runge-Kutta step of the CUDA path against the CPU oracle (C/OpenMP stencil
backend, bit-identical to the NumPy one -- tests/test_oracle_ckernels.py) on identical inputs.

Tolerances (BASELINE.json north_star): relative L2 per step <= 1e-12 (Float64), <= 1e-5 (Float32);
post-projection divergence at machine precision."""
import os

import numpy as np
import pytest

import oracle as O
from oracle import bc as obc, ckernels, initializers as oini, operators as oops, pressure as opr, timestep as ots

from gpu_helpers import make_setups, rel_l2, smooth_vec, uniform

pytestmark = pytest.mark.gpu

TOL = {np.float64: 1e-12, np.float32: 1e-5}


def ins():
    import insb200

    return insb200


@pytest.fixture(autouse=True)
def c_backend():
    """the threaded C restatement of the stencil operators (same bits as NumPy, much faster)"""
    cores = len(os.sched_getaffinity(0))
    old = oops.BACKEND, opr.WORKERS
    oops.BACKEND, opr.WORKERS = "c", cores
    ckernels.set_threads(cores)
    yield
    oops.BACKEND, opr.WORKERS = old


def tgv3d(dim, x, y, z):
    if dim == 0:
        return np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y) * np.sin(2 * np.pi * z) / 2
    if dim == 1:
        return -np.cos(2 * np.pi * x) * np.sin(2 * np.pi * y) * np.sin(2 * np.pi * z) / 2
    return 0 * x + 0 * y + 0 * z


def _step_both(so, sg, ps_o, ps_g, mname, st_o, st_g, f_o, f_g, p_g, dt, T, nsteps=1):
    I = ins()
    m_o = ots.LMWray3(T) if mname == "LMWray3" else getattr(ots, mname)(T)
    m_g = I.LMWray3(T) if mname == "LMWray3" else getattr(I.RKMethods, mname)(T)
    cache = ots.get_cache(m_o, st_o, so)
    sp_o = ots.create_stepper(m_o, so, ps_o, st_o, 0.0)
    sp_g = I.create_stepper(m_g, setup=sg, psolver=ps_g, state=st_g, t=0.0)
    for _ in range(nsteps):
        sp_o = ots.timestep_(m_o, f_o, sp_o, T(dt), cache)
        sp_g = I.timestep_(m_g, f_g, sp_g, T(dt), params=p_g)
    return sp_o, sp_g


def _max_div(I, sg, so, u_dev):
    return float(np.max(np.abs(I.to_host(I.divergence(u_dev, sg))[so.Ip])))


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("mname", ["RK44", "LMWray3", "RK33C2"])
def test_stage_kernel_many_tiles_and_chunks(T, mname):
    # 192 x 40 x 160: 3 x tiles + the one-column remainder strip, 5 y tiles, 2 z chunks (ZC = 128);
    # RK33C2 takes the general-tableau path (k storage), RK44 / LMWray3 the running-accumulator one
    I = ins()
    so, sg = make_setups((uniform(192), uniform(40, 0.5), uniform(160)), ("P", "P", "P"), T=T)
    ps_o, ps_g = opr.psolver_spectral(so), I.psolver_spectral(sg)
    u0 = opr.project(smooth_vec(so, 41), so, ps_o)
    obc.apply_bc_u(u0, 0.0, so)
    sp_o, sp_g = _step_both(so, sg, ps_o, ps_g, mname, {"u": u0.copy(order="F")}, {"u": I.to_device(sg, u0)},
                            ots.make_navierstokes(so, 2e-3), I.navierstokes_, dict(viscosity=2e-3), 2e-3, T)
    assert rel_l2(I.to_host(sp_g["state"]["u"]), sp_o["state"]["u"]) < TOL[T]
    scale = np.max(np.abs(u0)) * 192
    assert _max_div(I, sg, so, sp_g["state"]["u"]) < (1e-12 if T is np.float64 else 2e-4) * scale


@pytest.mark.parametrize("T", [np.float64, np.float32])
def test_stage_kernel_unaligned_odd_extent(T):
    # odd N0 (walls in x, 133 volumes): tiles start at cell 1, pairs move as two scalars; 3 x tiles, 2 z chunks
    I = ins()
    so, sg = make_setups((O.tanh_grid(0, 1, 133, 1.2), uniform(24), uniform(144)), ("D", "P", "P"), T=T)
    ps_o, ps_g = opr.psolver_tridiagonal(so), I.psolver_tridiagonal(sg)
    u0 = opr.project(smooth_vec(so, 42), so, ps_o)
    obc.apply_bc_u(u0, 0.0, so)
    sp_o, sp_g = _step_both(so, sg, ps_o, ps_g, "RK44", {"u": u0.copy(order="F")}, {"u": I.to_device(sg, u0)},
                            ots.make_navierstokes(so, 2e-3), I.navierstokes_, dict(viscosity=2e-3), 1e-3, T)
    got, ref = I.to_host(sp_g["state"]["u"]), sp_o["state"]["u"]
    for a in range(3):
        assert rel_l2(got[so.Iu[a] + (a,)], ref[so.Iu[a] + (a,)]) < TOL[T]


@pytest.mark.parametrize("T", [np.float64, np.float32])
def test_tgv3d_256_step(T):
    # BASELINE config 2 at half the extent (256^3; the 512^3 leg is bench.py --check): 4 x tiles, 2 z chunks
    I = ins()
    n = 256
    so, sg = make_setups((uniform(n), uniform(n), uniform(n)), ("P", "P", "P"), T=T)
    ps_o, ps_g = opr.psolver_spectral(so), I.psolver_spectral(sg)
    u0 = oini.velocityfield(so, tgv3d, psolver=ps_o)
    h = 1.0 / n
    dt = 0.5 * min(h * h / 1e-3 / 6, 2 * h)
    sp_o, sp_g = _step_both(so, sg, ps_o, ps_g, "RK44", {"u": u0.copy(order="F")}, {"u": I.to_device(sg, u0)},
                            ots.make_navierstokes(so, 1e-3), I.navierstokes_, dict(viscosity=1e-3), dt, T)
    assert rel_l2(I.to_host(sp_g["state"]["u"]), sp_o["state"]["u"]) < TOL[T]
    assert _max_div(I, sg, so, sp_g["state"]["u"]) < (1e-12 if T is np.float64 else 2e-4) * 0.5 * n


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("mname", ["RK44", "LMWray3"])
def test_rayleigh_benard_128_step(T, mname):
    # BASELINE config 4 at a quarter of the extent: 128 x 128 x 64, Boussinesq with dissipation,
    # psolver_transform (DCT in y and z), examples/RayleighBenard3D.jl:23-79
    I = ins()
    so, sg = make_setups((uniform(128, np.pi), uniform(128), uniform(64)), ("P", "D", "D"),
                         temp_bcs=("P", "S", (("D", 1.0), ("D", 0.0))), T=T)
    ps_o, ps_g = opr.psolver_transform(so), I.psolver_transform(sg)
    u0 = opr.project(0.05 * smooth_vec(so, 43), so, ps_o)
    obc.apply_bc_u(u0, 0.0, so)
    xp = so.xp
    temp0 = O.scalarfield(so)
    temp0[...] = (0.5 + np.sin(20 * xp[0])[:, None, None] * np.sin(20 * np.pi * xp[1])[None, :, None] / 100
                  + 0 * xp[2][None, None, :]).astype(T)
    obc.apply_bc_temp(temp0, 0.0, so)
    params = dict(viscosity=2.5e-4, conductivity=2.5e-4, gdir=2, gravity=1.0, dodissipation=True)
    sp_o, sp_g = _step_both(so, sg, ps_o, ps_g, mname, {"u": u0.copy(order="F"), "temp": temp0.copy(order="F")},
                            {"u": I.to_device(sg, u0), "temp": I.to_device(sg, temp0)}, ots.make_boussinesq(so, **params),
                            I.boussinesq_, params, 5e-3, T)
    got, ref = I.to_host(sp_g["state"]["u"]), sp_o["state"]["u"]
    for a in range(3):
        assert rel_l2(got[so.Iu[a] + (a,)], ref[so.Iu[a] + (a,)]) < TOL[T] * (1 if T is np.float64 else 2)
    assert rel_l2(I.to_host(sp_g["state"]["temp"])[so.Ip], sp_o["state"]["temp"][so.Ip]) < TOL[T]


def test_channel_256_step():
    # BASELINE config 5 at a quarter of the extent: 256 x 128 x 128 Float64, tanh-stretched walls in y,
    # constant body force, psolver_tridiagonal (examples/ChannelFlow.jl:153-176)
    I = ins()
    T = np.float64
    so, sg = make_setups((uniform(256, 4 * np.pi), O.tanh_grid(0, 2, 128, 1.4), uniform(128, 4 * np.pi / 3)),
                         ("P", "D", "P"), T=T)
    ps_o, ps_g = opr.psolver_tridiagonal(so), I.psolver_tridiagonal(sg)
    u0 = opr.project(smooth_vec(so, 44), so, ps_o)
    obc.apply_bc_u(u0, 0.0, so)
    nu = 1.0 / 180
    f_o = ots.make_navierstokes(so, nu, bodyforce=(1.0, 0.0, 0.0))
    sp_o, sp_g = _step_both(so, sg, ps_o, ps_g, "RK44", {"u": u0.copy(order="F")}, {"u": I.to_device(sg, u0)}, f_o,
                            I.navierstokes_, dict(viscosity=nu, bodyforce=(1.0, 0.0, 0.0)), 1e-3, T)
    got, ref = I.to_host(sp_g["state"]["u"]), sp_o["state"]["u"]
    for a in range(3):
        assert rel_l2(got[so.Iu[a] + (a,)], ref[so.Iu[a] + (a,)]) < TOL[T]


# ---- fused periodic path (csrc/stage_fused.cu): lazy projection + fused divergence + TMA plane staging ----
FUSED_GRIDS = [(32, 64, 32), (64, 32, 64), (128, 32, 256), (256, 64, 32)]


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("mname", ["RK44", "LMWray3"])
@pytest.mark.parametrize("n", FUSED_GRIDS)
def test_fused_step_matches_oracle_and_staged_path(n, mname, T):
    # power-of-two periodic boxes take the fused path: partial x tiles (32), several x / y tiles, two z
    # chunks (256), constant body force; compared with the oracle and with the stage-by-stage CUDA path
    I = ins()
    so, sg = make_setups((uniform(n[0]), uniform(n[1], 0.5), uniform(n[2], 2.0)), ("P", "P", "P"), T=T)
    ps_o, ps_g = opr.psolver_spectral(so), I.psolver_spectral(sg)
    u0 = opr.project(smooth_vec(so, 45), so, ps_o)
    obc.apply_bc_u(u0, 0.0, so)
    bf = (0.3, -0.2, 0.1)
    f_o = ots.make_navierstokes(so, 2e-3, bodyforce=bf)
    pg = dict(viscosity=2e-3, bodyforce=bf)
    sp_o, sp_g = _step_both(so, sg, ps_o, ps_g, mname, {"u": u0.copy(order="F")}, {"u": I.to_device(sg, u0)}, f_o,
                            I.navierstokes_, pg, 2e-3, T, nsteps=2)
    got = I.to_host(sp_g["state"]["u"])
    assert rel_l2(got, sp_o["state"]["u"]) < 2 * TOL[T]
    assert _max_div(I, sg, so, sp_g["state"]["u"]) < (1e-12 if T is np.float64 else 2e-4) * np.max(np.abs(u0)) * max(n)
    os.environ["INSB200_NO_FUSED_STEP"] = "1"
    try:
        _, sp_s = _step_both(so, sg, ps_o, ps_g, mname, {"u": u0.copy(order="F")}, {"u": I.to_device(sg, u0)}, f_o,
                             I.navierstokes_, pg, 2e-3, T, nsteps=2)
    finally:
        del os.environ["INSB200_NO_FUSED_STEP"]
    assert rel_l2(got, I.to_host(sp_s["state"]["u"])) < 2 * TOL[T]


def test_fused_path_is_taken():
    # the fused step launches 1 stencil + 5 FFT kernels per stage and one projection per step
    I = ins()
    so, sg = make_setups((uniform(64), uniform(64), uniform(64)), ("P", "P", "P"), T=np.float32)
    ps_g = I.psolver_spectral(sg)
    u = I.velocityfield(sg, tgv3d, psolver=ps_g)
    m = I.RKMethods.RK44(np.float32)
    sp = I.create_stepper(m, setup=sg, psolver=ps_g, state={"u": u}, t=0.0)
    sp = I.timestep_(m, I.navierstokes_, sp, np.float32(1e-3), params=dict(viscosity=1e-3))
    l0 = ps_g.plan.launch_count
    sp = I.timestep_(m, I.navierstokes_, sp, np.float32(1e-3), params=dict(viscosity=1e-3))
    assert ps_g.plan.launch_count - l0 == 4 * 6 + 1
