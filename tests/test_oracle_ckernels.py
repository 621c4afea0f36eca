"""The C/OpenMP restatement (oracle/c, the multithreaded CPU baseline of bench.py) gives the
same IEEE results as the NumPy oracle, operator by operator and over whole RK44 steps."""
import numpy as np
import pytest

import oracle as O
from oracle import bc as obc, ckernels, operators as oops, pressure as opr, timestep as ots

from gpu_helpers import uniform  # CPU-only helper despite the module name

CASES = [
    ("periodic3", (uniform(12), uniform(10, 2.0), uniform(14, 0.5)), ("P", "P", "P")),
    ("walls3", (O.tanh_grid(0, 1, 10, 1.2), O.tanh_grid(0, 1, 8, 1.1), O.cosine_grid(0, 1, 12)), ("D", "D", "D")),
    ("channel3", (uniform(12, 4.0), O.tanh_grid(0, 2, 10, 1.4), uniform(8, 2.0)), ("P", "D", "P")),
    ("periodic2", (uniform(16), uniform(24, 2.0)), ("P", "P")),
    ("walls2", (O.tanh_grid(0, 1, 16), O.tanh_grid(0, 1, 12, 1.3)), ("D", "D")),
]


@pytest.fixture(autouse=True)
def _numpy_backend():
    oops.BACKEND = "numpy"
    yield
    oops.BACKEND = "numpy"


def _setup(x, bcs, T):
    xs = tuple(np.asarray(v).astype(T) for v in x)

    def one(k):
        return {"P": O.PeriodicBC, "D": O.DirichletBC}[k]()

    return O.Setup(xs, {"u": tuple((one(k), one(k)) for k in bcs)})


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("name,x,bcs", CASES)
def test_c_operators_equal_numpy_oracle(name, x, bcs, T):
    assert ckernels.available()
    so = _setup(x, bcs, T)
    rng = np.random.default_rng(3)
    u = np.asfortranarray(rng.standard_normal(so.N + (so.D,)).astype(T))
    p = np.asfortranarray(rng.standard_normal(so.N).astype(T))
    res = {}
    for be in ("numpy", "c"):
        oops.BACKEND = be
        f = oops.convectiondiffusion_(O.vectorfield(so), u, so, 1e-2)
        c = oops.convection_(O.vectorfield(so), u, so)
        d = oops.diffusion_(O.vectorfield(so), u, so, 1e-2)
        div = oops.divergence_(O.scalarfield(so), u, so)
        sc = oops.scalewithvolume_(p.copy(order="F"), so)
        ap = oops.applypressure_(u.copy(order="F"), p, so)
        res[be] = (f, c, d, div, sc, ap)
    for a, b in zip(res["numpy"], res["c"]):
        assert np.array_equal(a, b), name


@pytest.mark.parametrize("T", [np.float64, np.float32])
def test_c_backend_rk44_step_equals_numpy_oracle(T):
    so = _setup((uniform(16), uniform(12), uniform(8)), ("P", "P", "P"), T)
    ps = opr.psolver_spectral(so)
    rng = np.random.default_rng(5)
    u0 = np.asfortranarray(rng.standard_normal(so.N + (3,)).astype(T))
    obc.apply_bc_u(u0, 0.0, so)
    u0 = opr.project(u0, so, ps)
    obc.apply_bc_u(u0, 0.0, so)
    out = {}
    for be in ("numpy", "c"):
        oops.BACKEND = be
        m = ots.RK44(T)
        st = {"u": u0.copy(order="F")}
        sp = ots.create_stepper(m, so, ps, st, 0.0)
        cache = ots.get_cache(m, st, so)
        for _ in range(2):
            sp = ots.timestep_(m, ots.make_navierstokes(so, 1e-2), sp, T(1e-3), cache)
        out[be] = sp["state"]["u"]
    assert np.array_equal(out["numpy"], out["c"])
