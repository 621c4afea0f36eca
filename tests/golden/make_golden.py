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


EXTRA_CASES = ("box3d_periodic_f64", "channel3d_f64", "box3d_periodic_f32")
LES_MODELS = (("smagorinsky", 0.17), ("wale", 0.569), ("vreman", 0.28), ("qr", 0.092))


def compute_extras(name, x, bcs, kind, fk, params, dt, T):
    """Rows N2-N4 of SURVEY.md §8f on the same grids: eddy-viscosity closure terms and an LES step,
    field diagnostics, energy spectrum (periodic boxes), psolver_cg."""
    from oracle import eddyviscosity as oev, postprocess as opp

    xs = tuple(np.asarray(v).astype(T) for v in x)
    so = O.Setup(xs, bcs)
    ps = getattr(opr, "psolver_" + kind)(so)
    rng = np.random.default_rng(sum(map(ord, name)) + 1)
    u = (0.5 * rng.standard_normal(so.N + (so.D,))).astype(T, order="F")
    obc.apply_bc_u(u, 0.0, so)
    u0 = opr.project(u, so, ps)
    obc.apply_bc_u(u0, 0.0, so)
    out = {"u0": u0}
    for model, C in LES_MODELS:
        f = np.zeros_like(u0)
        oev.eddy_viscosity_closure_(model, C, f, u0, so)
        out["les_" + model] = f
    force = ots.make_navierstokes_eddy(so, params["viscosity"], "wale", 0.569, (1.0, 0.0, 0.0))
    m = ots.RK44(T)
    st = {"u": u0.copy(order="F")}
    sp = ots.create_stepper(m, so, ps, st, 0.0)
    cache = ots.get_cache(m, st, so)
    for _ in range(2):
        sp = ots.timestep_(m, force, sp, T(dt), cache)
    out["les_wale_rk44_u"] = sp["state"]["u"]
    out["vorticity"] = opp.vorticity(u0, so)
    out["u_p"] = opp.interpolate_u_p(u0, so)
    out["vorticity_p"] = opp.interpolate_omega_p(out["vorticity"], so)
    out["qcrit"] = opp.qcrit(u0, so)
    out["ke"] = opp.kinetic_energy(u0, so)
    out["ke_first"] = opp.kinetic_energy(u0, so, True)
    out["total_ke"] = np.array([opp.total_kinetic_energy(u0, so), opp.total_kinetic_energy(u0, so, True)], dtype=np.float64)
    if kind == "spectral":
        out["spectrum"] = opp.energyspectrum(u0, so)[1]
    rhs = oops.scalewithvolume_(oops.divergence_(O.scalarfield(so), u, so), so)
    out["cg_rhs"] = rhs
    out["cg_sol"] = opr.psolver_cg(so, reltol=1e-13 if T is np.float64 else 1e-6)(rhs.copy(order="F"))
    return out


def main():
    for case in cases():
        out = compute(*case)
        np.savez_compressed(os.path.join(HERE, case[0] + ".npz"), **out)
        print(case[0], {k: v.shape for k, v in out.items()})
        if case[0] in EXTRA_CASES:
            out = compute_extras(*case)
            np.savez_compressed(os.path.join(HERE, case[0] + "_extras.npz"), **out)
            print(case[0] + "_extras", sorted(out))


if __name__ == "__main__":
    main()
