"""GPU parity of the on-device diagnostics against oracle/postprocess.py."""
import numpy as np
import pytest

import oracle as O
from oracle import postprocess as opp

from gpu_helpers import make_setups, random_vec, rel_l2, smooth_vec, uniform

pytestmark = pytest.mark.gpu

CASES = [
    ("periodic3", (uniform(16), uniform(12, 2.0), uniform(20, 0.5)), ("P", "P", "P")),
    ("walls3", (O.tanh_grid(0, 1, 12, 1.2), O.tanh_grid(0, 1, 10, 1.1), O.cosine_grid(0, 1, 14)), ("D", "D", "D")),
    ("channel3", (uniform(16, 4.0), O.tanh_grid(0, 2, 12, 1.4), uniform(8, 2.0)), ("P", "D", "P")),
    ("mixed3", (uniform(10), O.cosine_grid(0, 1, 8), uniform(12)), (("D", "O"), "S", "P")),
    ("periodic2", (uniform(16), uniform(24, 2.0)), ("P", "P")),
    ("walls2", (O.tanh_grid(0, 1, 16), O.tanh_grid(0, 1, 12, 1.3)), ("D", "D")),
]


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("name,x,bcs", CASES)
def test_diagnostics_match_oracle(name, x, bcs, T):
    import insb200 as I

    so, sg = make_setups(x, bcs, T=T)
    u = smooth_vec(so, 91)
    ud = I.to_device(sg, u)
    tol = 1e-13 if T is np.float64 else 2e-6
    # outputs start from the same random content: entries outside the reference's range stay untouched
    junk_v, junk_s = random_vec(so, 92), random_vec(so, 93)[..., 0].copy(order="F")
    vort_shape_vec = so.D == 3
    w_o = opp.vorticity_(junk_v.copy(order="F") if vort_shape_vec else junk_s.copy(order="F"), u, so)
    w_g = I.vorticity_(I.to_device(sg, junk_v if vort_shape_vec else junk_s), ud, sg)
    assert rel_l2(I.to_host(w_g), w_o) < tol
    up_o = opp.interpolate_u_p_(junk_v.copy(order="F"), u, so)
    up_g = I.interpolate_u_p_(I.to_device(sg, junk_v), ud, sg)
    assert rel_l2(I.to_host(up_g), up_o) < tol
    wp_o = opp.interpolate_omega_p_(junk_v.copy(order="F") if vort_shape_vec else junk_s.copy(order="F"), w_o, so)
    wp_g = I.interpolate_omega_p_(I.to_device(sg, junk_v if vort_shape_vec else junk_s), w_g, sg)
    assert rel_l2(I.to_host(wp_g), wp_o) < tol
    q_o = opp.qcrit_(junk_s.copy(order="F"), u, so)
    q_g = I.qcrit_(I.to_device(sg, junk_s), ud, sg)
    assert rel_l2(I.to_host(q_g), q_o) < (1e-12 if T is np.float64 else 2e-5)
    for first in (False, True):
        ke_o = opp.kinetic_energy_(junk_s.copy(order="F"), u, so, first)
        ke_g = I.kinetic_energy_(I.to_device(sg, junk_s), ud, sg, first)
        assert rel_l2(I.to_host(ke_g), ke_o) < tol
        tot_o = float(opp.total_kinetic_energy(u, so, first))
        tot_g = float(I.total_kinetic_energy(ud, sg, first))
        assert abs(tot_g - tot_o) <= (1e-13 if T is np.float64 else 2e-6) * abs(tot_o)


def test_diagnostics_at_bench_size_properties():
    """256^3 F32 Taylor-Green field: analytic vorticity / energy, finite Q (size-independent checks)."""
    import torch

    import insb200 as I

    n = 256
    T = np.float32
    x = tuple(np.linspace(0.0, 1.0, n + 1).astype(T) for _ in range(3))
    sg = I.Setup(x, {"u": ((I.PeriodicBC(), I.PeriodicBC()),) * 3})

    def U(dim, x, y, z):
        if dim == 0:
            return np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y) * np.sin(2 * np.pi * z) / 2
        if dim == 1:
            return -np.cos(2 * np.pi * x) * np.sin(2 * np.pi * y) * np.sin(2 * np.pi * z) / 2
        return 0 * x + 0 * y + 0 * z

    u = I.velocityfield(sg, U, doproject=False)
    e = float(I.total_kinetic_energy(u, sg))
    assert abs(e - 1.0 / 32) < 2e-4  # int (u^2 + v^2)/2 = 2 * (1/4) * (1/8) / 2
    w = I.vorticity(u, sg)
    q = I.qcrit(u, sg)
    assert bool(torch.isfinite(w).all()) and bool(torch.isfinite(q).all())
    # omega_z = dv/dx - du/dy = 2 pi sin(2 pi x) sin(2 pi y) sin(2 pi z): max = 2 pi
    assert abs(float(w[2].abs().max()) - 2 * np.pi) < 2e-2


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("n", [(64, 32, 128), (128, 64)])
def test_energy_spectrum_matches_oracle(n, T):
    import insb200 as I

    D = len(n)
    so, sg = make_setups(tuple(uniform(m, 2 * np.pi) for m in n), ("P",) * D, T=T)
    ps = I.psolver_spectral(sg)
    u = smooth_vec(so, 121) + 0.05 * random_vec(so, 122)
    kmax = int(np.sqrt(D) * max(n) / 2) + 1
    for km in (None, kmax):
        kap, want = opp.energyspectrum(u.astype(T), so, kmax=km)
        got = I.energyspectrum(I.to_device(sg, u), sg, ps, kmax=km)
        assert np.array_equal(got["kappa"], kap)
        scale = want.max()
        assert np.max(np.abs(got["ehat"].astype(np.float64) - want)) < (1e-12 if T is np.float64 else 2e-5) * scale
    # Parseval with all shells
    assert abs(float(got["ehat"].astype(np.float64).sum()) - 0.5 * np.mean(np.sum(u[so.Ip].astype(np.float64) ** 2, axis=-1))) < (
        1e-11 if T is np.float64 else 1e-4)
    # the Poisson solve still works afterwards (the workspace is shared)
    f = smooth_vec(so, 123)[..., 0].copy(order="F")
    a = I.to_host(ps(I.to_device(sg, f)))
    I.energyspectrum(I.to_device(sg, u), sg, ps)
    b = I.to_host(ps(I.to_device(sg, f)))
    assert np.array_equal(a, b)


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("n,L", [((64, 32, 128), (2 * np.pi, 1.0, 3.0)), ((128, 64), (1.0, 2.0))])
def test_turbulence_statistics_match_oracle(n, L, T):
    """turbulence_statistics (src/spectral.jl:144-257) through insb200_spectral_moments against the oracle's
    restatement, anisotropic box lengths so that every dimension's 2 pi / L enters."""
    import insb200 as I

    D = len(n)
    so, sg = make_setups(tuple(uniform(m, l) for m, l in zip(n, L)), ("P",) * D, T=T)
    ps = I.psolver_spectral(sg)
    u = smooth_vec(so, 131) + 0.05 * random_vec(so, 132)
    nu = 2e-3
    want = opp.turbulence_statistics(u.astype(T), so, nu)
    got = I.turbulence_statistics(I.to_device(sg, u), sg, nu, ps)
    assert set(got) == set(want)
    for k, v in want.items():
        assert abs(float(got[k]) - float(v)) <= (1e-11 if T is np.float64 else 2e-4) * abs(float(v)), (k, got[k], v)
    assert all(np.asarray(v).dtype == np.dtype(T) for v in got.values())
    # repeatable up to the order of the atomic adds, and the Poisson workspace is shared without harm
    again = I.turbulence_statistics(I.to_device(sg, u), sg, nu, ps)
    assert abs(float(again["diss"]) - float(got["diss"])) <= 1e-12 * float(got["diss"]) * (1 if T is np.float64 else 1e6)
    f = smooth_vec(so, 133)[..., 0].copy(order="F")
    a = I.to_host(ps(I.to_device(sg, f)))
    I.turbulence_statistics(I.to_device(sg, u), sg, nu, ps)
    assert np.array_equal(a, I.to_host(ps(I.to_device(sg, f))))
    with pytest.raises(AssertionError):
        so2, sg2 = make_setups(tuple(uniform(16) for _ in range(D)), ("D",) * D, T=T)
        I.turbulence_statistics(I.vectorfield(sg2), sg2, nu, I.default_psolver(sg2))


def test_field_observers_and_vtk_output(tmp_path):
    """observefield / observespectrum / vtk_writer (src/processors.jl:85-277) over the device diagnostics."""
    import xml.etree.ElementTree as ET

    import insb200 as I

    so, sg = make_setups((uniform(32, 2 * np.pi), uniform(32, 2 * np.pi), uniform(32, 2 * np.pi)), ("P", "P", "P"))
    ps = I.psolver_spectral(sg)
    u = opr_project(smooth_vec(so, 141), so)
    ud = I.to_device(sg, u)
    st = I.Observable({"u": ud, "t": 0.0, "n": 0})
    Ip = so.Ip
    want = {
        "velocity": opp.interpolate_u_p(u, so)[Ip],
        2: opp.interpolate_u_p(u, so)[Ip + (1,)],
        "velocitynorm": np.sqrt(2 * opp.kinetic_energy(u, so)[Ip]),
        "vorticity": opp.interpolate_omega_p(opp.vorticity(u, so), so)[Ip],
        "qcrit": opp.qcrit(u, so)[Ip],
    }
    for fn, w in want.items():
        f = I.observefield(st, setup=sg, fieldname=fn)
        assert f.value.shape == w.shape and rel_l2(f.value, w) < 1e-13, fn
    spec = I.observespectrum(st, setup=sg, psolver=ps)
    assert np.allclose(spec["ehat"].value, opp.energyspectrum(u, so)[1], rtol=0, atol=1e-12 * spec["ehat"].value.max())
    # observers follow the state
    f = I.observefield(st, setup=sg, fieldname="velocitynorm")
    st.value = {"u": 2 * ud, "t": 0.1, "n": 1}
    assert rel_l2(f.value, 2 * want["velocitynorm"]) < 1e-13
    assert np.allclose(spec["ehat"].value, 4 * opp.energyspectrum(u, so)[1], rtol=1e-12)
    # VTK output through solve_unsteady
    out, outputs = I.solve_unsteady(setup=sg, tlims=(0.0, 0.02), start={"u": ud}, method=I.RKMethods.RK44(), psolver=ps,
                                    dt=0.005, params=dict(viscosity=1e-2),
                                    processors={"vtk": I.vtk_writer(sg, nupdate=2, dir=str(tmp_path), filename="sol",
                                                                    fieldnames=("velocity", "qcrit"))})
    pvd = ET.parse(outputs["vtk"]).getroot()
    files = [d.attrib["file"] for d in pvd.iter("DataSet")]
    assert len(files) == 3  # n = 0, 2, 4
    root = ET.parse(str(tmp_path / files[-1])).getroot()
    arrays = {a.attrib["Name"]: a for a in root.iter("DataArray")}
    assert {"velocity", "qcrit", "TimeValue", "x", "y", "z"} <= set(arrays)
    vel = np.array(arrays["velocity"].text.split(), dtype=np.float64).reshape(-1, 3)
    up = I.to_host(I.interpolate_u_p(out["u"], sg))[Ip]
    assert np.allclose(vel[:, 0], up[..., 0].ravel(order="F")) and np.allclose(vel[:, 2], up[..., 2].ravel(order="F"))
    assert abs(float(arrays["TimeValue"].text) - 0.02) < 1e-12


def opr_project(u, so):
    from oracle import bc as obc, pressure as opr

    u = opr.project(u, so, opr.psolver_spectral(so))
    obc.apply_bc_u(u, 0.0, so)
    return u


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("n", [(32, 32, 32), (64, 32, 128), (64, 32)])
def test_random_field_on_device(n, T):
    """random_field (src/initializers.jl:109-183, test/spectral.jl:29-56) synthesised on the device with the
    solver's FFT passes: with the oracle's white noise injected the field equals the oracle's; with the device
    generator (Philox) it is divergence free on the staggered grid, its spectrum equals the target in every shell
    (the oracle's energyspectrum is the independent check), the same seed reproduces it (to round-off) and another one does not."""
    import insb200 as I
    from oracle import initializers as oini, operators as oops

    D = len(n)
    L = (1.0, 2.0, 0.5)[:D]
    so, sg = make_setups(tuple(uniform(m, l) for m, l in zip(n, L)), ("P",) * D, T=T)
    ps = I.psolver_spectral(sg)
    total, kpeak = 0.3, 5
    want = oini.random_field(so, 0.0, totalenergy=total, kpeak=kpeak, rng=np.random.default_rng(123))
    got = I.to_host(I.random_field(sg, 0.0, totalenergy=total, kpeak=kpeak, rng=np.random.default_rng(123), psolver=ps))
    assert rel_l2(got, want) < (1e-12 if T is np.float64 else 2e-5)
    kdiag = int(np.floor(np.sqrt(sum((m // 2) ** 2 for m in so.Np))))
    target = oini.orlandi_profile(np.arange(kdiag + 1, dtype=np.float64), kpeak=kpeak)
    target = total * target / target.sum()
    fields = []
    for seed in (7, 7, 8):
        ud = I.random_field(sg, 0.0, totalenergy=total, kpeak=kpeak, rng=seed, psolver=ps)
        u = I.to_host(ud).astype(np.float64)
        fields.append(u)
        so64 = make_setups(tuple(uniform(m, l) for m, l in zip(n, L)), ("P",) * D, T=np.float64)[0]
        div = oops.divergence_(O.scalarfield(so64), u, so64)
        umax = np.abs(u).max()
        assert np.max(np.abs(div[so64.Ip])) < (1e-10 if T is np.float64 else 2e-4) * umax * max(n)
        kap, e = opp.energyspectrum(u, so64, kmax=kdiag)
        assert abs(e.sum() - total) < (1e-12 if T is np.float64 else 1e-6)
        assert np.linalg.norm(e - target) <= (1e-10 if T is np.float64 else 1e-5) * np.linalg.norm(target)
        # ghost cells are the periodic images (apply_bc_u!)
        assert np.array_equal(u[0], u[-2]) and np.array_equal(u[-1], u[1])
    # same seed: the same noise bit for bit; the shell sums are accumulated with atomics, so the rescaling factors
    # (and the field) agree to round-off rather than bitwise
    assert rel_l2(fields[0], fields[1]) < (1e-13 if T is np.float64 else 1e-6)
    assert rel_l2(fields[0], fields[2]) > 0.5
