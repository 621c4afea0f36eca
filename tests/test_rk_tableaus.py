"""Butcher tableaus of the host mirror (insb200.RKMethods) and of the oracle against the reference's own
src/time_steppers/RKMethods.jl, extracted into tests/golden/rk_tableaus.json by tests/golden/make_rk_tableaus.py,
plus the order conditions the method names promise.  CPU only."""
import json
import os
from fractions import Fraction

import numpy as np
import pytest

from oracle import timestep as ots

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rk_tableaus.json")))


def _num(v):
    return float(Fraction(v)) if "/" in v else float(v)


def _gold(name):
    t = GOLD[name]
    return [[_num(v) for v in row] for row in t["A"]], [_num(v) for v in t["b"]], [_num(v) for v in t["c"]], _num(t["r"])


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("name", sorted(GOLD))
def test_host_mirror_tableaus_equal_the_reference(name, T):
    import insb200 as I

    A, b, c, r = _gold(name)
    want = ots.runge_kutta_method(A, b, c, r, T)  # methods.jl:71-89: convert to T, shift
    got = getattr(I.RKMethods, name)(T)
    tol = 0 if name != "HEM5" else 4 * np.finfo(T).eps
    for x, y in ((got.A, want.A), (got.b, want.b), (got.c, want.c)):
        assert x.dtype == np.dtype(T) and x.shape == y.shape
        assert np.max(np.abs(x.astype(np.float64) - y.astype(np.float64))) <= tol, name
    assert float(got.r) == float(want.r)


@pytest.mark.parametrize("name", ["FE11", "RK44", "Wray3", "RK33C2"])
def test_oracle_tableaus_equal_the_reference(name):
    A, b, c, r = _gold(name)
    want = ots.runge_kutta_method(A, b, c, r)
    got = getattr(ots, name)()
    assert np.array_equal(got.A, want.A) and np.array_equal(got.b, want.b) and np.array_equal(got.c, want.c)


ORDER = {"FE11": 1, "SSP22": 2, "SSP42": 2, "SSP33": 3, "SSP43": 3, "Wray3": 3, "RK56": 5, "DOPRI6": 5, "Mid22": 2,
         "MTE22": 2, "Heun33": 3, "RK33C2": 3, "RK33P2": 3, "RK44": 4, "RK44C2": 4, "RK44C23": 4, "RK44P2": 4,
         "HEM3": 3, "HEM3BS": 3, "HEM5": 4, "NSSP21": 1, "NSSP32": 2, "NSSP33": 3, "NSSP53": 3}


@pytest.mark.parametrize("name", sorted(GOLD))
def test_order_conditions(name):
    """Rooted-tree order conditions up to order 4 (and the quadrature conditions b.c^k = 1/(k+1) up to the order)."""
    A, b, c, _ = _gold(name)
    A, b, c = np.array(A), np.array(b), np.array(c)
    p = ORDER[name]
    tol = 1e-14
    assert abs(b.sum() - 1) < tol
    assert np.allclose(A.sum(axis=1), c, atol=tol) or name in ("RK33C2",)  # RK33C2's c is given explicitly and equals the row sums too
    for k in range(1, p):
        assert abs(b @ c**k - 1 / (k + 1)) < tol, (name, k)
    if p >= 3:
        assert abs(b @ A @ c - 1 / 6) < tol
    if p >= 4:
        assert abs(b @ (c * (A @ c)) - 1 / 8) < tol
        assert abs(b @ A @ c**2 - 1 / 12) < tol
        assert abs(b @ A @ A @ c - 1 / 24) < tol
