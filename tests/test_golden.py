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
