"""Golden vectors (tests/golden/*.npz, written by tests/golden/make_golden.py).

CPU: the oracle still reproduces the committed vectors (regression pin of the checker).
GPU: the CUDA path reproduces them through the C ABI, without executing oracle code for the
expected values.  Tolerances: relative L2 <= 1e-12 (Float64) / 1e-5 (Float32) per
BASELINE.json north_star; single operators with cancellation get 5e-13 / 3e-5."""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)

CASES = list(mg.cases())
IDS = [c[0] for c in CASES]


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def load(name):
    return dict(np.load(os.path.join(HERE, "golden", name + ".npz")))


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_oracle_reproduces_golden(case):
    want = load(case[0])
    got = mg.compute(*case)
    T = case[-1]
    assert set(got) == set(want)
    for k in want:
        assert rel(got[k], want[k]) <= (1e-12 if T is np.float64 else 2e-6), k


def _to_insb(bcs):
    import insb200 as I
    import oracle as O

    m = {O.PeriodicBC: I.PeriodicBC, O.DirichletBC: I.DirichletBC, O.SymmetricBC: I.SymmetricBC,
         O.PressureBC: I.PressureBC}

    def conv(b):
        return I.DirichletBC(b.u) if isinstance(b, O.DirichletBC) else m[type(b)]()

    return {k: tuple((conv(l), conv(r)) for l, r in v) for k, v in bcs.items()}


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_cuda_path_reproduces_golden(case):
    import insb200 as I

    name, x, bcs, kind, fk, params, dt, T = case
    g = load(name)
    tol = 1e-12 if T is np.float64 else 1e-5
    optol = 5e-13 if T is np.float64 else 3e-5
    sg = I.Setup(tuple(np.asarray(v).astype(T) for v in x), _to_insb(bcs))
    ps = getattr(I, "psolver_" + kind)(sg)
    # projection of the raw field
    u = I.to_device(sg, g["u_raw"])
    assert rel(I.to_host(I.divergence(u, sg)), g["div_u"]) <= optol
    u0 = I.project(u, sg, psolver=ps)
    I.apply_bc_u_(u0, 0.0, sg)
    assert rel(I.to_host(u0), g["u0"]) <= tol
    # Poisson solve
    p = ps(I.to_device(sg, g["p_rhs"]))
    inner = sg.slices(sg.Ip)
    assert rel(I.to_host(p)[inner], g["p_sol"][inner]) <= (1e-11 if T is np.float64 else 1e-4)
    # right-hand side and two steps of each stepper
    state = {"u": I.to_device(sg, g["u0"])}
    if fk == "bouss":
        state["temp"] = I.to_device(sg, g["temp0"])
        force, fp = I.boussinesq_, params
    else:
        force, fp = I.navierstokes_, params
    k = {kk: I.vectorfield(sg) if kk == "u" else I.scalarfield(sg) for kk in state}
    force(k, state, 0.0, setup=sg, **fp)
    assert rel(I.to_host(k["u"]), g["k_u"]) <= optol
    if fk == "bouss":
        assert rel(I.to_host(k["temp"]), g["k_temp"]) <= optol
    for mname, m in (("rk44", I.RKMethods.RK44(T)), ("lmwray3", I.LMWray3(T))):
        st = I.create_stepper(m, setup=sg, psolver=ps, state={kk: v.clone() for kk, v in state.items()}, t=0.0)
        for _ in range(2):
            st = I.timestep_(m, force, st, T(dt), params=fp)
        assert rel(I.to_host(st["state"]["u"]), g[mname + "_u"]) <= tol, mname
        if fk == "bouss":
            assert rel(I.to_host(st["state"]["temp"]), g[mname + "_temp"]) <= tol, mname


EXTRA = [c for c in CASES if c[0] in mg.EXTRA_CASES]


@pytest.mark.parametrize("case", EXTRA, ids=[c[0] for c in EXTRA])
def test_oracle_reproduces_golden_extras(case):
    want = load(case[0] + "_extras")
    got = mg.compute_extras(*case)
    T = case[-1]
    assert set(got) == set(want)
    for k in want:
        lim = 1e-12 if T is np.float64 else 2e-6
        if k == "cg_sol":  # iterative: converged to reltol, round-off decides the last iterations
            lim = 1e-9 if T is np.float64 else 1e-3
        assert rel(got[k], want[k]) <= lim, k


@pytest.mark.gpu
@pytest.mark.parametrize("case", EXTRA, ids=[c[0] for c in EXTRA])
def test_cuda_path_reproduces_golden_extras(case):
    import insb200 as I

    name, x, bcs, kind, fk, params, dt, T = case
    g = load(name + "_extras")
    f64 = T is np.float64
    tol = 1e-12 if f64 else 1e-5
    sg = I.Setup(tuple(np.asarray(v).astype(T) for v in x), _to_insb(bcs))
    ps = getattr(I, "psolver_" + kind)(sg)
    u0 = I.to_device(sg, g["u0"])
    models = {"smagorinsky": I.Smagorinsky, "wale": I.WALE, "vreman": I.Vreman, "qr": I.QR}
    for model, C in mg.LES_MODELS:
        f = I.vectorfield(sg)
        I.eddy_viscosity_closure_(models[model](C), f, u0, None, sg)
        assert rel(I.to_host(f), g["les_" + model]) <= (1e-11 if f64 else 3e-4), model
    m = I.RKMethods.RK44(T)
    st = I.create_stepper(m, setup=sg, psolver=ps, state={"u": u0.clone()}, t=0.0)
    fp = dict(viscosity=params["viscosity"], bodyforce=(1.0, 0.0, 0.0), eddyviscosity=I.WALE(0.569))
    for _ in range(2):
        st = I.timestep_(m, I.navierstokes_, st, T(dt), params=fp)
    assert rel(I.to_host(st["state"]["u"]), g["les_wale_rk44_u"]) <= tol
    w = I.vorticity(u0, sg)
    assert rel(I.to_host(w), g["vorticity"]) <= tol
    assert rel(I.to_host(I.interpolate_u_p(u0, sg)), g["u_p"]) <= tol
    assert rel(I.to_host(I.interpolate_omega_p(w, sg)), g["vorticity_p"]) <= tol
    assert rel(I.to_host(I.qcrit(u0, sg)), g["qcrit"]) <= (1e-12 if f64 else 2e-5)
    assert rel(I.to_host(I.kinetic_energy(u0, sg)), g["ke"]) <= tol
    assert rel(I.to_host(I.kinetic_energy(u0, sg, True)), g["ke_first"]) <= tol
    tk = [float(I.total_kinetic_energy(u0, sg)), float(I.total_kinetic_energy(u0, sg, True))]
    assert rel(tk, g["total_ke"]) <= (1e-13 if f64 else 2e-6)
    if kind == "spectral":
        if all(n >= 32 and n & (n - 1) == 0 for n in sg.Np):  # register-FFT path (what energyspectrum runs on)
            got = I.energyspectrum(u0, sg, ps)["ehat"]
            assert np.max(np.abs(got.astype(np.float64) - g["spectrum"])) <= (1e-12 if f64 else 2e-5) * g["spectrum"].max()
        else:
            with pytest.raises(I.Insb200Error):
                I.energyspectrum(u0, sg, ps)
    cg = I.psolver_cg(sg, reltol=1e-13 if f64 else 1e-6)
    sol = I.to_host(cg(I.to_device(sg, g["cg_rhs"])))
    inner = sg.slices(sg.Ip)
    a, b = sol[inner] - sol[inner].mean(), g["cg_sol"][inner] - g["cg_sol"][inner].mean()
    assert rel(a, b) <= (1e-8 if f64 else 5e-3)
