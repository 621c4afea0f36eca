"""GPU: every alternative kernel variant against the default variant (same library,
switched with the INSB200_* environment variables the launchers read at call time), plus the
host-buffer ensemble entry point against device-resident stepping.  The oracle comparisons of
the default variants live in test_gpu_parity.py."""
import contextlib
import ctypes as C
import os

import numpy as np
import pytest

import oracle as O
from oracle import bc as obc, pressure as opr

from gpu_helpers import make_setups, random_scal, rel_l2, smooth_vec, uniform

pytestmark = pytest.mark.gpu


def ins():
    import insb200

    return insb200


@contextlib.contextmanager
def env(**kw):
    old = {k: os.environ.get(k) for k in kw}
    try:
        for k, v in kw.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(v)
        yield
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


FUSE_CASES = [
    ("periodic3", (uniform(32), uniform(64, 2.0), uniform(32, 0.5)), ("P", "P", "P"), "spectral"),
    ("periodic2", (uniform(64), uniform(32)), ("P", "P"), "spectral"),
    ("rb3", (uniform(32, np.pi), uniform(64), uniform(32)), ("P", "D", "D"), "transform"),
    ("channel3", (uniform(64, 4.0), O.tanh_grid(0, 2, 24, 1.4), uniform(32, 2.0)), ("P", "D", "P"), "tridiagonal"),
    ("channel2", (uniform(32, 4.0), O.cosine_grid(0, 1, 16)), ("P", "D"), "tridiagonal"),
    # extents where whole blocks of row pairs fit into a plane: shared-memory staged variant
    ("staged_512_16_8", (uniform(512), uniform(16, 2.0), uniform(8, 0.5)), ("P", "P", "P"), "spectral"),
    ("staged_256_32_8", (uniform(256), uniform(32), uniform(8)), ("P", "P", "P"), "spectral"),
    ("staged_1024_8_4", (uniform(1024), uniform(8), uniform(4)), ("P", "P", "P"), "spectral"),
    ("staged_64_64_64", (uniform(64), uniform(64), uniform(64)), ("P", "P", "P"), "spectral"),
    ("staged2_512_64", (uniform(512), uniform(64)), ("P", "P"), "spectral"),
    ("staged_rb_128_64_32", (uniform(128, np.pi), uniform(64), uniform(32)), ("P", "D", "D"), "transform"),
    ("staged_channel_128_64_32", (uniform(128, 4.0), O.tanh_grid(0, 2, 64, 1.4), uniform(32, 2.0)), ("P", "D", "P"), "tridiagonal"),
]


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("name,x,bcs,kind", FUSE_CASES)
def test_fused_divergence_matches_separate_pass(name, x, bcs, kind, T):
    """divergence! + scalewithvolume! evaluated inside the first FFT pass == separate kernel,
    and both reproduce the oracle's project!."""
    I = ins()
    so, sg = make_setups(x, bcs, T=T)
    mk_g = {"spectral": I.psolver_spectral, "transform": I.psolver_transform, "tridiagonal": I.psolver_tridiagonal}[kind]
    mk_o = {"spectral": opr.psolver_spectral, "transform": opr.psolver_transform, "tridiagonal": opr.psolver_tridiagonal}[kind]
    ps_g, ps_o = mk_g(sg), mk_o(so)
    u = smooth_vec(so, 31)
    want = opr.project(u.copy(order="F"), so, ps_o)
    out = {}
    for tag, fuse in (("fused", 2), ("fused_unstaged", 1), ("separate", None)):
        with env(INSB200_FUSED_DIVERGENCE=fuse):
            ud = I.to_device(sg, u)
            I.project_(ud, sg, psolver=ps_g, p=I.scalarfield(sg))
            out[tag] = I.to_host(ud)
    tol = 1e-12 if T is np.float64 else 1e-5
    inner = tuple(slice(1, -1) for _ in range(so.D))
    for tag in out:
        assert rel_l2(out[tag][inner], want[inner]) < tol, (name, tag)
    assert rel_l2(out["fused"], out["separate"]) < (1e-13 if T is np.float64 else 1e-6)
    assert rel_l2(out["fused_unstaged"], out["separate"]) < (1e-13 if T is np.float64 else 1e-6)


@pytest.mark.parametrize("T", [np.float64, np.float32])
def test_fused_divergence_inside_periodic_step(T):
    """The time stepper's variant (periodic wrap instead of ghost reads): staged == unstaged == separate."""
    I = ins()
    so, sg = make_setups((uniform(128), uniform(32), uniform(16)), ("P", "P", "P"), T=T)
    ps = I.psolver_spectral(sg)
    u0 = smooth_vec(so, 61)
    m = I.RKMethods.RK44(T)
    res = {}
    for tag, fuse in (("fused", 2), ("fused_unstaged", 1), ("separate", None)):
        with env(INSB200_FUSED_DIVERGENCE=fuse):
            sp = I.create_stepper(m, setup=sg, psolver=ps, state={"u": I.to_device(sg, u0)}, t=0.0)
            for _ in range(2):
                sp = I.timestep_(m, I.navierstokes_, sp, T(2e-3), params=dict(viscosity=1e-3))
            res[tag] = I.to_host(sp["state"]["u"])
    tol = 1e-13 if T is np.float64 else 2e-6
    assert rel_l2(res["fused"], res["separate"]) < tol
    assert rel_l2(res["fused_unstaged"], res["separate"]) < tol


@pytest.mark.parametrize("T", [np.float64, np.float32])
def test_boussinesq_split_matches_single_kernel(T):
    """marching momentum kernel + temperature/buoyancy pass == the one-kernel direct variant."""
    I = ins()
    x = (uniform(32, np.pi), uniform(24), uniform(20))
    bcs = ("P", "D", "D")
    tb = ("P", "S", (("D", 1.0), ("D", 0.0)))
    so, sg = make_setups(x, bcs, temp_bcs=tb, T=T)
    ps_g = I.psolver_transform(sg)
    u0 = smooth_vec(so, 41)
    temp0 = obc.apply_bc_temp(random_scal(so, 42), 0.0, so)
    params = dict(viscosity=1e-2, conductivity=2e-2, gdir=2, gravity=1.0, dodissipation=True)
    res = {}
    for tag, val, tile in (("split", None, 1), ("split_notile", None, 0), ("direct", 1, None)):
        with env(INSB200_BOUSS_DIRECT=val, INSB200_BOUSS_TILE=tile):
            for mname in ("RK44", "Wray3"):
                m = getattr(I.RKMethods, mname)(T)
                st = {"u": I.to_device(sg, u0), "temp": I.to_device(sg, temp0)}
                sp = I.create_stepper(m, setup=sg, psolver=ps_g, state=st, t=0.0)
                for _ in range(2):
                    sp = I.timestep_(m, I.boussinesq_, sp, 5e-3, params=params)
                res[tag, mname] = (I.to_host(sp["state"]["u"]), I.to_host(sp["state"]["temp"]))
    tol = 1e-13 if T is np.float64 else 1e-6
    for mname in ("RK44", "Wray3"):
        for k in (0, 1):
            assert rel_l2(res["split", mname][k], res["direct", mname][k]) < tol, (mname, k)
            assert np.array_equal(res["split", mname][k], res["split_notile", mname][k]), (mname, k)


def test_host_ensemble_matches_device_stepping():
    """insb200_step_rk_host_ensemble (pipelined H2D / step / D2H over members) == stepping each
    member on the device."""
    import torch

    I = ins()
    from insb200._lib import check, lib
    from insb200.api import _coef_dtype, _params

    T = np.float32
    n = 64
    so, sg = make_setups((uniform(n), uniform(n), uniform(n)), ("P", "P", "P"), T=T)
    ps = I.psolver_spectral(sg)
    m = I.RKMethods.RK44(T)
    dt = T(1e-3)
    members = [smooth_vec(so, 50 + i) for i in range(5)]
    want = []
    for u0 in members:
        sp = I.create_stepper(m, setup=sg, psolver=ps, state={"u": I.to_device(sg, u0)}, t=0.0)
        for _ in range(2):
            sp = I.timestep_(m, I.navierstokes_, sp, dt, params=dict(viscosity=1e-3))
        want.append(I.to_host(sp["state"]["u"]))
    hosts = []
    for u0 in members:
        h = torch.empty(I.to_device(sg, u0).shape, dtype=torch.float32, pin_memory=True)
        h.copy_(I.to_device(sg, u0))
        hosts.append(h)
    s = len(m.b)
    A = (C.c_double * (s * s))(*[float(v) for v in np.asarray(m.A).reshape(-1)])
    cc = (C.c_double * s)(*[float(v) for v in m.c])
    pr = _params(sg, viscosity=1e-3)
    ptrs = (C.c_void_p * len(hosts))(*[h.data_ptr() for h in hosts])
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    check(lib().insb200_step_rk_host_ensemble(ps.plan.h, A, cc, s, _coef_dtype(m.T, dt), ptrs, None, len(hosts), 0.0,
                                              float(dt), 2, C.byref(pr), st))
    for h, w in zip(hosts, want):
        assert np.array_equal(I.to_host(h.cuda()), w)


@pytest.mark.parametrize("mname", ["RK44", "LMWray3"])
@pytest.mark.parametrize("adaptive", [False, True])
def test_native_time_loop_matches_python_loop(mname, adaptive):
    """insb200_solve_unsteady (loop inside the library, CUDA-graph replay for fixed dt) ==
    one library call per step driven from Python; processors see the same states."""
    I = ins()
    T = np.float64
    so, sg = make_setups((uniform(32), uniform(32)), ("P", "P"), T=T)
    ps = I.psolver_spectral(sg)
    u0 = I.to_device(sg, smooth_vec(so, 71))
    I.project_(u0, sg, psolver=ps, p=I.scalarfield(sg))
    I.apply_bc_u_(u0, 0.0, sg)
    m = I.LMWray3(T) if mname == "LMWray3" else I.RKMethods.RK44(T)
    kw = dict(setup=sg, tlims=(0.0, 0.1), start={"u": u0}, method=m, psolver=ps, params=dict(viscosity=1e-2))
    kw.update(dict(cfl=0.5, n_adapt_dt=3) if adaptive else dict(dt=0.004))
    res = {}
    for tag, opts in (("python", dict(native=False)), ("native", dict(native=True, graph=False)),
                      ("graph", dict(native=True, graph=True))):
        log = []
        out, outputs = I.solve_unsteady(processors={"log": I.timelogger(sg, nupdate=5, out=log),
                                                    "save": I.fieldsaver(sg, nupdate=10)}, **kw, **opts)
        res[tag] = (I.to_host(out["u"]), float(out["t"]), [(e["n"], e["t"], e["umax"]) for e in log],
                    [(s["n"], s["u"]) for s in outputs["save"]])
    for tag in ("native", "graph"):
        assert np.array_equal(res[tag][0], res["python"][0]), tag
        assert res[tag][1] == res["python"][1]
        assert res[tag][2] == res["python"][2], tag
        assert len(res[tag][3]) == len(res["python"][3])
        for (na, ua), (nb, ub) in zip(res[tag][3], res["python"][3]):
            assert na == nb and np.array_equal(ua, ub)
    assert abs(res["python"][1] - 0.1) < 1e-12


def test_native_time_loop_stops_on_bad_dt():
    I = ins()
    so, sg = make_setups((uniform(16), uniform(16)), ("P", "P"))
    ps = I.psolver_spectral(sg)
    u0 = I.vectorfield(sg)  # u = 0 and nu = 1e6: the diffusive limit h^2 / (4 nu) is below 1e-10
    with pytest.warns(UserWarning, match="out of bounds"):
        out, _ = I.solve_unsteady(setup=sg, tlims=(0.0, 1.0), start={"u": u0}, psolver=ps, params=dict(viscosity=1e8))
    assert float(out["t"]) == 0.0


@pytest.mark.parametrize("T", [np.float64, np.float32])
def test_first_stage_reuses_ring_planes_bit_identical(T):
    """First stage of a step (vbase = acc_in = u): operands taken from the shared-memory plane
    ring == fetched separately."""
    I = ins()
    so, sg = make_setups((uniform(70), uniform(24), uniform(20)), ("P", "P", "P"), T=T)
    ps = I.psolver_spectral(sg)
    u0 = smooth_vec(so, 81)
    res = {}
    for tag, val in (("ring", None), ("fetched", 1)):
        with env(INSB200_STAGE_NO_SAME=val):
            for mname in ("RK44", "LMWray3", "FE11"):
                m = I.LMWray3(T) if mname == "LMWray3" else getattr(I.RKMethods, mname)(T)
                sp = I.create_stepper(m, setup=sg, psolver=ps, state={"u": I.to_device(sg, u0)}, t=0.0)
                for _ in range(2):
                    sp = I.timestep_(m, I.navierstokes_, sp, T(1e-3), params=dict(viscosity=1e-2))
                res[tag, mname] = I.to_host(sp["state"]["u"])
    for mname in ("RK44", "LMWray3", "FE11"):
        assert np.array_equal(res["ring", mname], res["fetched", mname]), mname


def test_native_loop_graph_replay_with_large_shared_memory_kernels():
    """CUDA-graph capture of a step whose FFT passes need > 48 KB of dynamic shared memory
    (the launchers raise the limit with cudaFuncSetAttribute on every call, also while capturing)."""
    I = ins()
    T = np.float32
    so, sg = make_setups((uniform(512), uniform(32), uniform(1024)), ("P", "P", "P"), T=T)
    ps = I.psolver_spectral(sg)
    u0 = I.to_device(sg, smooth_vec(so, 131))
    I.project_(u0, sg, psolver=ps, p=I.scalarfield(sg))
    I.apply_bc_u_(u0, 0.0, sg)
    kw = dict(setup=sg, tlims=(0.0, 1e-4), start={"u": u0}, method=I.RKMethods.RK44(T), psolver=ps, dt=2e-5,
              params=dict(viscosity=1e-3))
    a, _ = I.solve_unsteady(native=False, **kw)
    b, _ = I.solve_unsteady(native=True, graph=True, **kw)
    ua = I.to_host(a["u"])
    assert np.all(np.isfinite(ua)) and np.array_equal(ua, I.to_host(b["u"]))


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("bcs", [("P", "P", "P"), ("P", "D", "P")])
def test_user_defined_force_composed_stage_by_stage(T, bcs):
    """A force the library does not know (src/solver.jl:55-61: any `force!` with the built-in signature) makes
    timestep_ compose the step on the host: force -> insb200_stage_combine_project per stage, the statestart
    update of LMWray3 through insb200_axpy.  With the built-in right-hand side wrapped in a user function the
    result must equal the fused one-call steppers (general tableaus, diagonal tableaus, low-storage Wray3) and,
    through them, the oracle; the stage arrays are allocated once and no torch kernel runs inside a step."""
    import torch

    I = ins()
    so, sg = make_setups((uniform(40), uniform(24), uniform(16)), bcs, T=T)
    ps = I.psolver_spectral(sg) if bcs[1] == "P" else I.psolver_tridiagonal(sg)
    u0 = smooth_vec(so, 17)
    calls = []

    def my_force(force, state, t, *, setup, cache=None, **params):
        calls.append(float(t))
        return I.navierstokes_(force, state, t, setup=setup, **params)

    dt = T(2e-3)
    tol = 1e-13 if T is np.float64 else 2e-6
    for mname in ("RK44", "RK33C2", "Wray3", "FE11", "LMWray3"):
        m = I.LMWray3(T) if mname == "LMWray3" else getattr(I.RKMethods, mname)(T)
        res = {}
        for tag, f in (("fused", I.navierstokes_), ("staged", my_force)):
            sp = I.create_stepper(m, setup=sg, psolver=ps, state={"u": I.to_device(sg, u0)}, t=0.0)
            I.apply_bc_u_(sp["state"]["u"], 0.0, sg)
            for _ in range(2):
                sp = I.timestep_(m, f, sp, dt, params=dict(viscosity=1e-2))
            res[tag] = I.to_host(sp["state"]["u"])
            assert abs(sp["t"] - 2 * float(dt)) < 1e-12
        assert rel_l2(res["staged"], res["fused"]) <= tol, mname
    assert len(calls) == 2 * (4 + 3 + 3 + 1 + 3)
    # the oracle, once, through the staged path (RK44, one step)
    m = I.RKMethods.RK44(T)
    sp = I.create_stepper(m, setup=sg, psolver=ps, state={"u": I.to_device(sg, u0)}, t=0.0)
    I.apply_bc_u_(sp["state"]["u"], 0.0, sg)
    sp = I.timestep_(m, my_force, sp, dt, params=dict(viscosity=1e-2))
    import oracle.timestep as ots
    from oracle import bc as obc2

    uo = obc2.apply_bc_u(u0.copy(order="F"), 0.0, so)
    pso = opr.psolver_spectral(so) if bcs[1] == "P" else opr.psolver_tridiagonal(so)
    m_o, st_o = ots.RK44(T), {"u": uo}
    sp_o = ots.timestep_(m_o, ots.make_navierstokes(so, 1e-2), ots.create_stepper(m_o, so, pso, st_o, 0.0), dt,
                         ots.get_cache(m_o, st_o, so))
    want = sp_o["state"]["u"]
    assert rel_l2(I.to_host(sp["state"]["u"]), want) <= (1e-12 if T is np.float64 else 1e-5)
    # no torch kernels inside a staged step: only the library's launches (and no allocation: same cache arrays)
    cache = sg._plans["stage_cache"][("u",)]
    ptrs = [k["u"].data_ptr() for k in cache["k"]]
    try:
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            sp = I.timestep_(m, my_force, sp, dt, params=dict(viscosity=1e-2))
            torch.cuda.synchronize()
        names = [e.key for e in prof.key_averages() if getattr(e, "device_type", None) is not None]
        kernels = [n for n in names if "insb" in n]
        foreign = [n for n in names if ("at::" in n or "elementwise" in n or "vectorized" in n)]
        if kernels:  # (CUPTI available)
            assert not foreign, foreign
    except (ImportError, RuntimeError):
        pass
    assert [k["u"].data_ptr() for k in sg._plans["stage_cache"][("u",)]["k"]] == ptrs


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("coef64", [False, True])
def test_stage_combine_and_axpy_round_like_the_reference_broadcast(T, coef64):
    """insb200_stage_combine / insb200_axpy: u = u_start + sum_j c_j k_j evaluated like `@. u += dt * A[i, j] * k`
    (step_explicit_runge_kutta.jl:28-32) -- every product and every sum rounded separately in
    promote(field type, coefficient type), stored in the field type; bit-exact against NumPy."""
    import ctypes as C

    import torch

    from insb200 import _lib
    from insb200._lib import check, lib
    from insb200.api import _base_plan, _ptr, _stream

    I = ins()
    so, sg = make_setups((uniform(12), uniform(10), uniform(8)), ("P", "P", "P"), T=T)
    rng = np.random.default_rng(5)
    h = _base_plan(sg).h
    shape = (3,) + tuple(reversed(sg.N))
    base = rng.standard_normal(shape).astype(T)
    ks = [rng.standard_normal(shape).astype(T) for _ in range(4)]
    CT = np.float64 if (coef64 or T is np.float64) else np.float32
    cdt = _lib.F64 if coef64 else (_lib.F64 if T is np.float64 else _lib.F32)
    coefs = [float(CT(v)) for v in (0.5e-3, -0.25e-3, 1.7e-3, 0.1e-3)]
    want = base.astype(CT)
    for c, k in zip(coefs, ks):
        want = (want.astype(T).astype(CT) + CT(c) * k.astype(CT))  # rounded to the field type after every axpy
    want = want.astype(T)
    dev = [torch.from_numpy(k).cuda() for k in ks]
    b = torch.from_numpy(base).cuda()
    out = torch.empty_like(b)
    kp = (C.c_void_p * 4)(*[_ptr(d) for d in dev])
    cf = (C.c_double * 4)(*coefs)
    check(lib().insb200_stage_combine(h, _ptr(out), _ptr(b), 4, kp, cf, cdt, 3, _stream()))
    assert np.array_equal(out.cpu().numpy(), want)
    # in place (dst aliases base) and axpy
    check(lib().insb200_stage_combine(h, _ptr(b), _ptr(b), 4, kp, cf, cdt, 3, _stream()))
    assert np.array_equal(b.cpu().numpy(), want)
    y = torch.from_numpy(base).cuda()
    check(lib().insb200_axpy(h, _ptr(y), coefs[0], _ptr(dev[0]), cdt, 3, _stream()))
    assert np.array_equal(y.cpu().numpy(), (base.astype(CT) + CT(coefs[0]) * ks[0].astype(CT)).astype(T))
    # scalar fields (ncomp = 1) and the copy (nk = 0)
    s_out = torch.empty(tuple(reversed(sg.N)), dtype=b.dtype, device="cuda")
    check(lib().insb200_stage_combine(h, _ptr(s_out), _ptr(b), 0, None, None, cdt, 1, _stream()))
    assert np.array_equal(s_out.cpu().numpy(), want[0])
    # more than 8 stage arrays are refused
    many = (C.c_void_p * 9)(*[_ptr(dev[0])] * 9)
    assert lib().insb200_stage_combine(h, _ptr(out), _ptr(b), 9, many, (C.c_double * 9)(*[0.0] * 9), cdt, 3, _stream()) != 0


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("mname", ["RK44", "LMWray3"])
def test_chained_steps_of_the_native_loop_are_bit_identical(mname, T):
    """Inside insb200_solve_unsteady the fused periodic path keeps the state lazy between steps (unprojected velocity +
    pressure: no explicit projection into u, the next step's first stage projects on load) and materialises u only for
    call-backs, CFL evaluations and at the end.  The result must equal stepping from Python -- one library call per
    step, u materialised every time -- bit for bit: fixed dt with and without CUDA-graph replay, with processors that
    look at the state every third step, and adaptive dt with a CFL evaluation every third step."""
    I = ins()
    so, sg = make_setups((uniform(70), uniform(24), uniform(20)), ("P", "P", "P"), T=T)
    ps = I.psolver_spectral(sg)
    u0 = I.to_device(sg, smooth_vec(so, 33))
    I.project_(u0, sg, psolver=ps, p=I.scalarfield(sg))
    I.apply_bc_u_(u0, 0.0, sg)
    m = I.LMWray3(T) if mname == "LMWray3" else I.RKMethods.RK44(T)
    base = dict(setup=sg, start={"u": u0}, method=m, psolver=ps, params=dict(viscosity=1e-2))
    for name, kw, procs in (("fixed", dict(tlims=(0.0, 0.014), dt=0.002), False),
                            ("fixed+processors", dict(tlims=(0.0, 0.014), dt=0.002), True),
                            ("adaptive", dict(tlims=(0.0, 0.02), cfl=0.4, n_adapt_dt=3), False)):
        res = {}
        for tag, opts in (("python", dict(native=False)), ("native", dict(native=True, graph=False)),
                          ("graph", dict(native=True, graph=True)), ("unchained", dict(native=True, graph=False))):
            log = []
            pr = {"log": I.timelogger(sg, nupdate=3, out=log), "save": I.fieldsaver(sg, nupdate=3)} if procs else {}
            with env(INSB200_NO_CHAINED_STEPS=1 if tag == "unchained" else None):
                out, outputs = I.solve_unsteady(processors=pr, **base, **kw, **opts)
            res[tag] = (I.to_host(out["u"]), float(out["t"]), [(e["n"], e["umax"]) for e in log],
                        [(s["n"], s["u"]) for s in outputs.get("save", [])])
        for tag in ("native", "graph", "unchained"):
            assert np.array_equal(res[tag][0], res["python"][0]), (name, tag)
            assert res[tag][1] == res["python"][1] and res[tag][2] == res["python"][2], (name, tag)
            for (na, ua), (nb, ub) in zip(res[tag][3], res["python"][3]):
                assert na == nb and np.array_equal(ua, ub), (name, tag)
        assert np.all(np.isfinite(res["python"][0]))
