#!/usr/bin/env python
"""Generates tests/golden/*.npz: small input/output vectors of the hot path.

The reference itself cannot be executed here (no Julia in the image), so these
vectors come from the CPU oracle (oracle/, a line-by-line restatement of the reference) AFTER
the oracle passed the reference's own property / manufactured-solution tests
(tests/test_oracle_*.py).  They pin the oracle against regressions and give the CUDA path a
fixed target that does not depend on the oracle's code at test time.  Bit-level parity with a
Julia build remains unpinned (DESIGN.md "Oracle").

    python tests/golden/make_golden.py          # rewrites the fixtures
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle as O  # noqa: E402
from oracle import bc as obc, initializers as oini, operators as oops, pressure as opr, timestep as ots  # noqa: E402

P, D_, S_ = O.PeriodicBC, O.DirichletBC, O.SymmetricBC


def cases():
    """name -> (x, bc dict, psolver kind, force kind, params, dt)"""
    u16 = np.linspace(0.0, 1.0, 17)
    yield ("tgv2d_periodic_f64", (np.linspace(0, 2 * np.pi, 17),) * 2, {"u": ((P(), P()),) * 2}, "spectral", "ns",
           dict(viscosity=5e-4), 0.02, np.float64)
    yield ("box3d_periodic_f64", (u16[:9] * 2, u16[:13], u16), {"u": ((P(), P()),) * 3}, "spectral", "ns",
           dict(viscosity=1e-3), 1e-3, np.float64)
    yield ("box3d_periodic_f32", (u16, u16, u16), {"u": ((P(), P()),) * 3}, "spectral", "ns",
           dict(viscosity=1e-3), 1e-3, np.float32)
    yield ("channel3d_f64", (np.linspace(0, 4.0, 9), O.tanh_grid(0, 2, 12, 1.4), np.linspace(0, 2.0, 9)),
           {"u": ((P(), P()), (D_(), D_()), (P(), P()))}, "tridiagonal", "ns", dict(viscosity=1e-2), 1e-3, np.float64)
    yield ("rb3d_f64", (np.linspace(0, np.pi, 9), np.linspace(0, 1, 11), np.linspace(0, 1, 9)),
           {"u": ((P(), P()), (D_(), D_()), (D_(), D_())),
            "temp": ((P(), P()), (S_(), S_()), (D_(1.0), D_(0.0)))}, "transform", "bouss",
           dict(viscosity=2.5e-3, conductivity=2.5e-3, gravity=1.0, gdir=2, dodissipation=True), 1e-3, np.float64)


def compute(name, x, bcs, kind, fk, params, dt, T):
    if True:
        xs = tuple(np.asarray(v).astype(T) for v in x)
        so = O.Setup(xs, bcs)
        ps = getattr(opr, "psolver_" + kind)(so)
        rng = np.random.default_rng(sum(map(ord, name)))
        u = (0.5 * rng.standard_normal(so.N + (so.D,))).astype(T, order="F")
        obc.apply_bc_u(u, 0.0, so)
        u0 = opr.project(u, so, ps)
        obc.apply_bc_u(u0, 0.0, so)
        out = {"u_raw": u, "u0": u0}
        state = {"u": u0.copy(order="F")}
        if fk == "bouss":
            temp = rng.standard_normal(so.N).astype(T, order="F")
            obc.apply_bc_temp(temp, 0.0, so)
            state["temp"] = temp
            out["temp0"] = temp.copy(order="F")
            force = ots.make_boussinesq(so, **params)
        else:
            force = ots.make_navierstokes(so, params["viscosity"])
        # right-hand side
        k = {kk: np.zeros_like(v) for kk, v in state.items()}
        force(k, state, 0.0)
        out["k_u"] = k["u"]
        if fk == "bouss":
            out["k_temp"] = k["temp"]
        # divergence, Poisson solve of a random right-hand side
        out["div_u"] = oops.divergence_(O.scalarfield(so), u, so)
        rhs = rng.standard_normal(so.N).astype(T, order="F")
        out["p_rhs"] = rhs
        out["p_sol"] = ps(rhs.copy(order="F"))
        for mname, m in (("rk44", ots.RK44(T)), ("lmwray3", ots.LMWray3(T))):
            st = {kk: v.copy(order="F") for kk, v in state.items()}
            sp = ots.create_stepper(m, so, ps, st, 0.0)
            cache = ots.get_cache(m, st, so)
            for _ in range(2):
                sp = ots.timestep_(m, force, sp, T(dt), cache)
            out[mname + "_u"] = sp["state"]["u"]
            if fk == "bouss":
                out[mname + "_temp"] = sp["state"]["temp"]
        return out


def main():
    for case in cases():
        out = compute(*case)
        np.savez_compressed(os.path.join(HERE, case[0] + ".npz"), **out)
        print(case[0], {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
