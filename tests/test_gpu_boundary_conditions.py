"""GPU: the reference's boundary-condition tests (test/boundary_conditions.jl, the stage-time test of
test/timesteppers.jl:44-91) through the C ABI -- function-valued DirichletBC (host-evaluated planes,
insb200_set_bc_u_plane), PressureBC with its double ghost volume, SymmetricBC, psolver_cg as the
default solver -- each also compared with the oracle run of the same case."""
import numpy as np
import pytest

import oracle as O
from oracle import bc as obc, operators as oops, pressure as opr, timestep as ots

from gpu_helpers import rel_l2

pytestmark = pytest.mark.gpu


def _pair(x, bcs_o, bcs_g):
    import insb200 as I

    xs = tuple(np.asarray(v, dtype=np.float64) for v in x)
    return O.Setup(xs, {"u": bcs_o}), I.Setup(xs, {"u": bcs_g})


def test_poiseuille_flow_with_pressure_outlet():
    import insb200 as I

    uin = lambda dim, x, y, t: (6 * y * (1 - y) if dim == 0 else 0 * y) + 0 * x
    so, sg = _pair((np.linspace(0.0, 4.0, 33), np.linspace(0.0, 1.0, 17)),
                   ((O.DirichletBC(uin), O.PressureBC()), (O.DirichletBC(), O.DirichletBC())),
                   ((I.DirichletBC(uin), I.PressureBC()), (I.DirichletBC(), I.DirichletBC())))
    y = so.xu[0][1].reshape(1, -1)
    u0 = O.vectorfield(so)
    u0[:, :, 0] = 6 * y * (1 - y)
    obc.apply_bc_u(u0, 0.0, so)
    ud = I.to_device(sg, u0)
    I.apply_bc_u_(ud, 0.0, sg)
    assert np.array_equal(I.to_host(ud), u0)  # the function planes land where the reference puts them
    ps_g = I.default_psolver(sg)
    assert ps_g.kind == "cg"
    ps_g = I.psolver_cg(sg, reltol=1e-12)
    out_g, _ = I.solve_unsteady(setup=sg, tlims=(0.0, 0.5), start={"u": ud}, psolver=ps_g, dt=0.01,
                                params=dict(viscosity=0.05))
    out_o = ots.solve_unsteady(so, (0.0, 0.5), {"u": u0}, ots.make_navierstokes(so, 0.05),
                               psolver=opr.psolver_cg(so, reltol=1e-12), dt=0.01)
    u = I.to_host(out_g["u"])
    assert rel_l2(u, out_o["u"]) < 1e-9
    jr = slice(*sg.Ip[1])
    flux = lambda i: np.sum(u[i, jr, 0] * so.Delta[1][jr])
    is_ = range(so.Iu[0][0].start, so.Iu[0][0].stop)
    fl = [flux(i) for i in [is_[0] - 1] + list(is_[::8]) + [is_[-1]]]
    assert max(fl) - min(fl) < 1e-8
    assert np.max(np.abs(I.to_host(I.divergence(out_g["u"], sg))[so.Ip])) < 1e-9


def test_time_derivative_planes_and_stage_times():
    """dudt = true ghost fill (:44-62) and the pulsating plug flow (timesteppers.jl:44-91): a Dirichlet
    function of time makes timestep_ compose the stages on the host with planes refreshed per stage."""
    import insb200 as I

    uin = lambda dim, x, y, t: ((1 + 0.5 * np.sin(2 * np.pi * t)) if dim == 0 else 0.0) + 0 * x + 0 * y
    so, sg = _pair((np.linspace(0.0, 4.0, 17), np.linspace(0.0, 1.0, 9)),
                   ((O.DirichletBC(uin), O.PressureBC()), (O.SymmetricBC(), O.SymmetricBC())),
                   ((I.DirichletBC(uin), I.PressureBC()), (I.SymmetricBC(), I.SymmetricBC())))
    z = O.vectorfield(so)
    for t in (0.13, 0.71):
        want = obc.apply_bc_u(z.copy(order="F"), t, so, dudt=True)
        got = I.to_host(I.apply_bc_u_(I.to_device(sg, z), t, sg, dudt=True))
        assert np.allclose(got, want, rtol=0, atol=1e-12)
        i1 = so.Iu[0][0].start - 1
        assert np.allclose(got[i1, so.Ip[1], 0], np.pi * np.cos(2 * np.pi * t), rtol=1e-6)
    u0 = O.vectorfield(so)
    u0[..., 0] = 1.0
    obc.apply_bc_u(u0, 0.0, so)
    dt, tend = 0.005, 0.1
    for m_o, m_g in ((ots.LMWray3(), I.LMWray3()), (ots.RK44(), I.RKMethods.RK44())):
        out_o = ots.solve_unsteady(so, (0.0, tend), {"u": u0}, ots.make_navierstokes(so, 0.05), method=m_o,
                                   psolver=opr.psolver_cg(so, reltol=1e-13), dt=dt)
        out_g, _ = I.solve_unsteady(setup=sg, tlims=(0.0, tend), start={"u": I.to_device(sg, u0)}, method=m_g,
                                    psolver=I.psolver_cg(sg, reltol=1e-13), dt=dt, params=dict(viscosity=0.05))
        u = I.to_host(out_g["u"])
        assert rel_l2(u, out_o["u"]) < 1e-10
        assert np.max(np.abs(u[so.Iu[0] + (0,)] - uin(0, 0.0, 0.0, tend))) < 1e-10
        assert np.max(np.abs(I.to_host(I.divergence(out_g["u"], sg))[so.Ip])) < 1e-10


def test_uniform_flow_with_lateral_pressure_boundaries():
    import insb200 as I

    so, sg = _pair((np.linspace(0.0, 2.0, 33), np.linspace(0.0, 1.0, 17)),
                   ((O.DirichletBC((1.0, 0.0)), O.PressureBC()), (O.PressureBC(), O.PressureBC())),
                   ((I.DirichletBC((1.0, 0.0)), I.PressureBC()), (I.PressureBC(), I.PressureBC())))
    u0 = O.vectorfield(so)
    u0[:, :, 0] = 1.0
    obc.apply_bc_u(u0, 0.0, so)
    out_g, _ = I.solve_unsteady(setup=sg, tlims=(0.0, 0.2), start={"u": I.to_device(sg, u0)},
                                psolver=I.psolver_cg(sg, reltol=1e-13), dt=0.005, params=dict(viscosity=0.05))
    u = I.to_host(out_g["u"])
    assert np.max(np.abs(u[so.Iu[0] + (0,)] - 1)) < 1e-11
    assert np.max(np.abs(u[so.Iu[1] + (1,)])) < 1e-11
    assert np.max(np.abs(I.to_host(I.divergence(out_g["u"], sg))[so.Ip])) < 1e-11
