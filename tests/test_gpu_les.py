"""GPU parity of the eddy-viscosity closures against oracle/eddyviscosity.py:
the closure term alone, and RK44 / LMWray3 steps with navierstokes! + closure (the force_eddy!
of examples/ChannelFlow.jl)."""
import numpy as np
import pytest

import oracle as O
from oracle import bc as obc, eddyviscosity as oev, pressure as opr, timestep as ots

from gpu_helpers import make_setups, rel_l2, smooth_vec, uniform

pytestmark = pytest.mark.gpu

CASES = [
    ("periodic", (uniform(16), uniform(12, 2.0), uniform(20, 0.5)), ("P", "P", "P"), "spectral"),
    ("channel", (uniform(16, 4.0), O.tanh_grid(0, 2, 12, 1.4), uniform(8, 2.0)), ("P", "D", "P"), "tridiagonal"),
    ("walls", (O.tanh_grid(0, 1, 12, 1.2), O.tanh_grid(0, 1, 10, 1.1), O.cosine_grid(0, 1, 14)), ("D", "D", "D"), None),
    ("mixed", (uniform(10), O.cosine_grid(0, 1, 8), uniform(12)), ("S", "D", "P"), None),
    # two x blocks and three y blocks of the 64 x 4 launch tiles, a different BC on either side
    ("wide", (uniform(70, 3.0), O.tanh_grid(0, 1, 9, 1.2), O.cosine_grid(0, 1, 6)), ("P", ("S", "D"), ("D", "S")), None),
    # two z chunks of the tile kernels (64 planes per block), partial tiles in x and y
    ("tall", (uniform(37, 2.0), O.tanh_grid(0, 1, 11, 1.2), uniform(70, 3.0)), ("S", "D", "P"), None),
    # PressureBC side: two ghost layers, stays on the four-pass path (les_needs_tensor)
    ("outflow", (O.tanh_grid(0, 1, 12, 1.2), uniform(8), uniform(10)), (("D", "O"), "D", "P"), None),
]
MODELS = [("smagorinsky", "Smagorinsky", 0.17), ("wale", "WALE", 0.569), ("vreman", "Vreman", 0.28), ("qr", "QR", 0.092)]


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("mo,mg,C", MODELS)
@pytest.mark.parametrize("name,x,bcs,kind", CASES)
def test_closure_term_matches_oracle(name, x, bcs, kind, mo, mg, C, T):
    import insb200 as I

    so, sg = make_setups(x, bcs, T=T)
    u = smooth_vec(so, 101)
    f0 = smooth_vec(so, 102)
    want = oev.eddy_viscosity_closure_(mo, C, f0.copy(order="F"), u, so)
    fd = I.to_device(sg, f0)
    I.eddy_viscosity_closure_(getattr(I, mg)(C), fd, I.to_device(sg, u), None, sg)
    got = I.to_host(fd)
    # compare the closure term itself (f0 is O(1), the term can be small)
    term_o, term_g = want - f0, got - f0
    scale = max(np.linalg.norm(term_o.ravel()), 1e-300)
    err = np.linalg.norm((term_g - term_o).ravel()) / scale
    assert err < (1e-11 if T is np.float64 else 3e-4), (name, mo, err)


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("mo,mg,C", MODELS)
@pytest.mark.parametrize("name,x,bcs,kind", CASES[:6])
def test_fused_closure_equals_four_pass_closure(name, x, bcs, kind, mo, mg, C, T, monkeypatch):
    """Default path (two fused kernels, nothing but nu_t stored: shared-memory tile kernels + the boundary shell;
    INSB200_LES_TILE=0: one thread per cell) against the four-pass path with the stored gradient / stress tensors
    (INSB200_LES_UNFUSED): same arithmetic up to the reciprocal metric tables."""
    import insb200 as I

    so, sg = make_setups(x, bcs, T=T)
    u = I.to_device(sg, smooth_vec(so, 201))
    f0 = smooth_vec(so, 202)
    out = []
    for env in ({}, {"INSB200_LES_TILE": "0"}, {"INSB200_LES_UNFUSED": "1"}):
        monkeypatch.delenv("INSB200_LES_UNFUSED", raising=False)
        monkeypatch.delenv("INSB200_LES_TILE", raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        fd = I.to_device(sg, f0)
        I.eddy_viscosity_closure_(getattr(I, mg)(C), fd, u, None, sg)
        out.append(I.to_host(fd) - f0)
    scale = max(np.linalg.norm(out[2].ravel()), 1e-300)
    for o in out[:2]:
        err = np.linalg.norm((o - out[2]).ravel()) / scale
        assert err < (1e-12 if T is np.float64 else 1e-5), (name, mo, err)
    # the same cell functions, whatever feeds them (up to the compiler's contraction choices per instantiation)
    assert np.linalg.norm((out[0] - out[1]).ravel()) / scale < (1e-14 if T is np.float64 else 1e-6)
    assert scale > 1e-6


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("name,x,bcs,kind", CASES[:2])
def test_les_steps_match_oracle(name, x, bcs, kind, T):
    import insb200 as I

    so, sg = make_setups(x, bcs, T=T)
    mk_o = {"spectral": opr.psolver_spectral, "tridiagonal": opr.psolver_tridiagonal}[kind]
    mk_g = {"spectral": I.psolver_spectral, "tridiagonal": I.psolver_tridiagonal}[kind]
    ps_o, ps_g = mk_o(so), mk_g(sg)
    u0 = opr.project(smooth_vec(so, 111), so, ps_o)
    obc.apply_bc_u(u0, 0.0, so)
    bodyforce = (1.0, 0.0, 0.0)
    for (mo, mg, C), mname in zip(MODELS, ("RK44", "LMWray3", "Wray3", "RK44")):
        f_o = ots.make_navierstokes_eddy(so, 1e-2, mo, C, bodyforce)
        m_o = ots.LMWray3(T) if mname == "LMWray3" else getattr(ots, mname)(T)
        m_g = I.LMWray3(T) if mname == "LMWray3" else getattr(I.RKMethods, mname)(T)
        st_o = {"u": u0.copy(order="F")}
        sp_o = ots.create_stepper(m_o, so, ps_o, st_o, 0.0)
        cache = ots.get_cache(m_o, st_o, so)
        sp_g = I.create_stepper(m_g, setup=sg, psolver=ps_g, state={"u": I.to_device(sg, u0)}, t=0.0)
        params = dict(viscosity=1e-2, bodyforce=bodyforce, eddyviscosity=getattr(I, mg)(C))
        for step in range(2):
            sp_o = ots.timestep_(m_o, f_o, sp_o, T(5e-3), cache)
            sp_g = I.timestep_(m_g, I.navierstokes_, sp_g, T(5e-3), params=params)
            e = rel_l2(I.to_host(sp_g["state"]["u"]), sp_o["state"]["u"])
            assert e < (1e-12 if T is np.float64 else 1e-5) * (step + 1), (name, mo, mname, step, e)
