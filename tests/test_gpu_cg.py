"""GPU parity of psolver_cg against the oracle's CG and the manufactured
solution of test/psolvers.jl:13-36, on boundary conditions / grids no FFT-type solver covers."""
import numpy as np
import pytest

import oracle as O
from oracle import bc as obc, operators as oops, pressure as opr, timestep as ots

from gpu_helpers import make_setups, random_scal, rel_l2, smooth_vec, uniform

pytestmark = pytest.mark.gpu

CASES = [
    ("cavity3", (O.tanh_grid(0, 1, 12, 1.2), O.tanh_grid(0, 1, 10, 1.1), O.cosine_grid(0, 1, 14)), ("D", "D", "D")),
    ("cavity2", (O.cosine_grid(0, 1, 24), O.cosine_grid(0, 1, 20)), ("D", "D")),
    ("outflow3", (uniform(10), O.cosine_grid(0, 1, 8), uniform(12)), (("D", "O"), "S", "P")),
    ("outflow2", (uniform(16, 4.0), uniform(12)), (("D", "O"), "S")),
    ("periodic3", (uniform(12), uniform(10, 2.0), uniform(8, 0.5)), ("P", "P", "P")),
    ("stretched_periodic2", (uniform(16), O.tanh_grid(0, 1, 16, 1.3)), ("P", "D")),
]


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("name,x,bcs", CASES)
def test_cg_matches_oracle_and_manufactured_solution(name, x, bcs, T):
    import insb200 as I

    so, sg = make_setups(x, bcs, T=T)
    ps_o, ps_g = opr.psolver_cg(so), I.psolver_cg(sg)
    assert ps_g.kind == "cg"
    pe = O.scalarfield(so)
    pe[so.Ip] = np.random.default_rng(3).standard_normal(so.Np).astype(T)
    obc.apply_bc_p(pe, 0.0, so)
    lap = oops.laplacian_(O.scalarfield(so), pe, so)
    want = ps_o(lap.copy(order="F"))
    got = I.to_host(ps_g(I.to_device(sg, lap)))
    has_pressure_bc = any("O" in (b if isinstance(b, tuple) else (b,)) for b in bcs)

    def fix(a):  # the solution is defined up to a constant unless a PressureBC pins it
        return a[so.Ip] if has_pressure_bc else a[so.Ip] - a[so.Ip].mean()

    # both are converged to reltol = sqrt(eps) in the residual: compare at that level
    lvl = 2e-5 if T is np.float64 else 2e-2
    assert rel_l2(fix(got), fix(want)) < lvl, name
    # (Float32: reltol = sqrt(eps) = 3e-4 in the residual of an ill-conditioned stretched-grid operator)
    assert rel_l2(fix(got), fix(pe)) < (lvl if T is np.float64 else 0.1), name
    # same stopping rule: iteration counts agree up to the read-back interval (4) and round-off
    assert abs(ps_g.iterations - ps_o.iterations) <= 8, (ps_g.iterations, ps_o.iterations)
    # ghosts as the reference's CG leaves them: apply_bc_p! of the solution, 0 at Dirichlet sides
    chk = got.copy(order="F")
    obc.apply_bc_p(chk, 0.0, so)
    assert np.array_equal(chk, got)


def test_lid_driven_cavity_steps_with_default_psolver():
    """All-wall stretched cavity (examples/LidDrivenCavity*.jl): default_psolver -> psolver_cg;
    two RK44 steps against the oracle stepping with its own CG, tight CG tolerance on both."""
    import insb200 as I

    T = np.float64
    x = (O.cosine_grid(0, 1, 16), O.cosine_grid(0, 1, 12), O.cosine_grid(0, 1, 10))
    lid = (("D", (0.0, 0.0, 0.0)), ("D", (1.0, 0.0, 0.2)))
    so, sg = make_setups(x, ("D", lid, "D"), T=T)
    ps_o = opr.psolver_cg(so, reltol=1e-13)
    ps_g = I.psolver_cg(sg, reltol=1e-13)
    assert I.default_psolver(sg).kind == "cg"
    u0 = O.vectorfield(so)
    obc.apply_bc_u(u0, 0.0, so)
    m_o, m_g = ots.RK44(T), I.RKMethods.RK44(T)
    st_o = {"u": u0.copy(order="F")}
    sp_o = ots.create_stepper(m_o, so, ps_o, st_o, 0.0)
    cache = ots.get_cache(m_o, st_o, so)
    sp_g = I.create_stepper(m_g, setup=sg, psolver=ps_g, state={"u": I.to_device(sg, u0)}, t=0.0)
    for _ in range(2):
        sp_o = ots.timestep_(m_o, ots.make_navierstokes(so, 1e-2), sp_o, 1e-2, cache)
        sp_g = I.timestep_(m_g, I.navierstokes_, sp_g, 1e-2, params=dict(viscosity=1e-2))
    assert rel_l2(I.to_host(sp_g["state"]["u"]), sp_o["state"]["u"]) < 1e-9
    div = I.to_host(I.divergence(sp_g["state"]["u"], sg))[so.Ip]
    assert np.max(np.abs(div)) < 1e-9
