"""Host-side logic of the Python mirror that needs no GPU: the allocating (non-mutating) variants of the reference
API must leave their arguments alone and compose the in-place entry points the way the reference does
(src/solver.jl:9-13,38-48, src/pressure.jl:16, src/boundary_conditions.jl:148-160, time_steppers/step.jl:17), checked
with the in-place functions replaced by CPU stand-ins; spectral_stuff against the oracle's independent shell binning."""
import numpy as np
import pytest
import torch

import oracle as O
from oracle import postprocess as opp

from gpu_helpers import make_setups, smooth_vec, uniform


@pytest.fixture
def api(monkeypatch):
    import insb200.api as A

    calls = []
    monkeypatch.setattr(A, "apply_bc_u_", lambda u, t, setup, dudt=False, plan=None: (calls.append(("bcu", t, dudt)), u.add_(1))[1])
    monkeypatch.setattr(A, "apply_bc_p_", lambda p, t, setup, plan=None: (calls.append(("bcp", t)), p.add_(2))[1])
    monkeypatch.setattr(A, "apply_bc_temp_", lambda x, t, setup, plan=None: (calls.append(("bct", t)), x.add_(3))[1])
    monkeypatch.setattr(A, "scalewithvolume_", lambda p, setup: p.mul_(2))
    monkeypatch.setattr(A, "convection_", lambda f, u, setup: f.add_(10 * u))
    monkeypatch.setattr(A, "diffusion_", lambda f, u, setup, visc: f.add_(visc * u))

    def boussinesq_(force, state, t, *, setup, cache=None, viscosity, conductivity, gdir, gravity, dodissipation):
        force["u"].add_(state["u"] * viscosity + gravity)
        force["temp"].add_(state["temp"] * conductivity + (1 if dodissipation else 0))
        return force

    monkeypatch.setattr(A, "boussinesq_", boussinesq_)

    def timestep_(method, force, stepper, dt, *, params=None, ode_cache=None, force_cache=None):
        for v in stepper["state"].values():
            v.add_(dt)
        return A.create_stepper(method, setup=stepper["setup"], psolver=stepper["psolver"], state=stepper["state"],
                                t=stepper["t"] + dt, n=stepper["n"] + 1)

    monkeypatch.setattr(A, "timestep_", timestep_)
    A._calls = calls
    return A


def test_allocating_variants_leave_their_arguments_alone(api):
    A = api
    u, p = torch.arange(6.0).reshape(1, 2, 3), torch.ones(2, 3)
    u0, p0 = u.clone(), p.clone()
    assert torch.equal(A.apply_bc_u(u, 0.5, None, dudt=True), u0 + 1) and torch.equal(u, u0)
    assert A._calls[-1] == ("bcu", 0.5, True)
    assert torch.equal(A.apply_bc_p(p, 0.0, None), p0 + 2) and torch.equal(p, p0)
    assert torch.equal(A.apply_bc_temp(p, 1.0, None), p0 + 3) and torch.equal(p, p0)
    assert torch.equal(A.scalewithvolume(p, None), 2 * p0) and torch.equal(p, p0)
    # poisson(psolver, f) = poisson!(psolver, copy(f))
    solved = A.poisson(lambda f: f.mul_(-1), p)
    assert torch.equal(solved, -p0) and torch.equal(p, p0)
    # navierstokes = convection + diffusion, summed afterwards; boussinesq through the fused right-hand side
    f = A.navierstokes({"u": u}, 0.0, setup=None, viscosity=0.25)
    assert torch.equal(f["u"], 10 * u0 + 0.25 * u0) and torch.equal(u, u0)
    st = {"u": u, "temp": p}
    f = A.boussinesq(st, 0.0, setup=None, viscosity=2.0, conductivity=3.0, gdir=2, gravity=0.5, dodissipation=True)
    assert torch.equal(f["u"], 2 * u0 + 0.5) and torch.equal(f["temp"], 3 * p0 + 1)
    assert torch.equal(u, u0) and torch.equal(p, p0)
    assert A.get_cache(A.navierstokes_, None) is None


def test_timestep_returns_a_new_state(api):
    A = api
    u = torch.zeros(3)
    sp = A.create_stepper("m", setup="s", psolver="p", state={"u": u}, t=1.0, n=4)
    new = A.timestep("m", A.navierstokes_, sp, 0.5, params=dict(viscosity=1.0))
    assert torch.equal(u, torch.zeros(3)) and sp["t"] == 1.0 and sp["n"] == 4
    assert torch.equal(new["state"]["u"], torch.full((3,), 0.5)) and new["t"] == 1.5 and new["n"] == 5
    assert new["state"]["u"].data_ptr() != u.data_ptr()


@pytest.mark.parametrize("n", [(16, 8, 12), (12, 10), (9, 8)])
def test_spectral_stuff_reproduces_the_oracle_spectrum(n):
    import insb200.api as A

    D = len(n)
    so, sg = make_setups(tuple(uniform(m, 2 * np.pi) for m in n), ("P",) * D)
    u = smooth_vec(so, 7)
    stuff = A.spectral_stuff(sg)
    assert stuff["Nhat"] == (n[0] // 2 + 1,) + tuple(n[1:])
    assert np.array_equal(stuff["kappa"], np.arange(min(n) // 2 + 1))
    ef = opp.spectral_energy_field(u, so).reshape(-1, order="F")
    ehat = np.array([ef[i].sum() for i in stuff["energyinds"]])
    kap, want = opp.energyspectrum(u, so)
    assert np.allclose(ehat, want, rtol=1e-12, atol=1e-14 * want.max())
    # every stored mode sits in at most one shell, shells are disjoint
    allinds = np.concatenate(stuff["inds"])
    assert len(np.unique(allinds)) == len(allinds)
