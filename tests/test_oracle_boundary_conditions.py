"""test/boundary_conditions.jl re-run on the oracle (known-answer pins for the Dirichlet-function,
Pressure and Symmetric boundary handling and for psolver_cg, which stands in for the reference's
direct solver on these setups)."""
import numpy as np

import oracle as O
from oracle import bc as obc, operators as oops, pressure as opr, timestep as ots


def test_poiseuille_flow_with_pressure_outlet():
    # test/boundary_conditions.jl:1-42 (integrated to t = 2 instead of 20: the start is the steady state)
    uin = lambda dim, x, y, t: (6 * y * (1 - y) if dim == 0 else 0 * y) + 0 * x
    bcs = {"u": ((O.DirichletBC(uin), O.PressureBC()), (O.DirichletBC(), O.DirichletBC()))}
    so = O.Setup((np.linspace(0.0, 4.0, 33), np.linspace(0.0, 1.0, 17)), bcs)
    y = so.xu[0][1].reshape(1, -1)
    u0 = O.vectorfield(so)
    u0[:, :, 0] = 6 * y * (1 - y)
    obc.apply_bc_u(u0, 0.0, so)
    ps = opr.psolver_cg(so, reltol=1e-12)
    out = ots.solve_unsteady(so, (0.0, 2.0), {"u": u0}, ots.make_navierstokes(so, 0.05), psolver=ps, dt=0.01)
    u = out["u"]
    jr = so.Ip[1]
    flux = lambda i: np.sum(u[i, jr, 0] * so.Delta[1][jr])
    is_ = range(so.Iu[0][0].start, so.Iu[0][0].stop)
    stations = [is_[0] - 1] + list(is_[::8]) + [is_[-1]]
    fl = [flux(i) for i in stations]
    assert max(fl) - min(fl) < 1e-8
    assert np.max(np.abs(oops.divergence_(O.scalarfield(so), u, so)[so.Ip])) < 1e-9
    imid = is_[len(is_) // 2]
    want = 6 * so.xu[0][1][jr] * (1 - so.xu[0][1][jr])
    # Julia's `≈` on vectors is norm-wise; measured 2.4e-3 (same as the reference's comment)
    assert np.linalg.norm(u[imid, jr, 0] - want) <= 5e-3 * np.linalg.norm(want)


def test_time_derivative_of_dirichlet_boundary_conditions():
    # test/boundary_conditions.jl:44-62
    uin = lambda dim, x, y, t: (1 + 0.5 * np.sin(2 * np.pi * t) if dim == 0 else 0.0) + 0 * x + 0 * y
    bcs = {"u": ((O.DirichletBC(uin), O.PressureBC()), (O.SymmetricBC(), O.SymmetricBC()))}
    so = O.Setup((np.linspace(0.0, 4.0, 17), np.linspace(0.0, 1.0, 9)), bcs)
    u = O.vectorfield(so)
    i1 = so.Iu[0][0].start - 1
    i2 = so.Iu[1][0].start - 1
    for t in (0.13, 0.71):
        d = obc.apply_bc_u(u.copy(order="F"), t, so, dudt=True)
        assert np.allclose(d[i1, so.Ip[1], 0], np.pi * np.cos(2 * np.pi * t), rtol=1e-6)
        assert np.all(d[i2, so.Ip[1], 1] == 0)


def test_uniform_flow_with_lateral_pressure_boundaries():
    # test/boundary_conditions.jl:64-84: the low-side PressureBC exercises the double ghost volume
    bcs = {"u": ((O.DirichletBC((1.0, 0.0)), O.PressureBC()), (O.PressureBC(), O.PressureBC()))}
    so = O.Setup((np.linspace(0.0, 2.0, 33), np.linspace(0.0, 1.0, 17)), bcs)
    u0 = O.vectorfield(so)
    u0[:, :, 0] = 1.0
    obc.apply_bc_u(u0, 0.0, so)
    ps = opr.psolver_cg(so, reltol=1e-13)
    out = ots.solve_unsteady(so, (0.0, 0.2), {"u": u0}, ots.make_navierstokes(so, 0.05), psolver=ps, dt=0.005)
    u = out["u"]
    assert np.max(np.abs(u[so.Iu[0] + (0,)] - 1)) < 1e-11
    assert np.max(np.abs(u[so.Iu[1] + (1,)])) < 1e-11
    assert np.max(np.abs(oops.divergence_(O.scalarfield(so), u, so)[so.Ip])) < 1e-11
