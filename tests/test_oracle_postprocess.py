"""Oracle diagnostics (oracle/postprocess.py) against analytic fields and identities."""
import numpy as np
import pytest

import oracle as O
from oracle import bc as obc, initializers as oini, postprocess as opp

from gpu_helpers import uniform


def _periodic(n, D, T=np.float64, L=2 * np.pi):
    x = tuple(np.linspace(0.0, L, n + 1).astype(T) for _ in range(D))
    return O.Setup(x, {"u": ((O.PeriodicBC(), O.PeriodicBC()),) * D})


def test_tgv2d_vorticity_energy_and_q():
    n = 128
    so = _periodic(n, 2)
    u = oini.velocityfield(so, lambda d, x, y: (-np.sin(x) * np.cos(y) if d == 0 else np.cos(x) * np.sin(y)) + 0 * x * y,
                           doproject=False)
    w = opp.vorticity(u, so)  # at the cell corners (right/top faces of cell I): omega = -2 sin x sin y
    N = so.N[0]
    xc, yc = so.x[0][1:N], so.x[1][1:N]
    want = -2 * np.sin(xc)[:, None] * np.sin(yc)[None, :]
    assert np.max(np.abs(w[: N - 1, : N - 1] - want)) < 2e-3
    # total kinetic energy of the 2-D Taylor-Green vortex on [0, 2pi]^2: pi^2 (both variants, O(h^2))
    assert abs(opp.total_kinetic_energy(u, so) - np.pi**2) < 2e-3
    assert abs(opp.total_kinetic_energy(u, so, interpolate_first=True) - np.pi**2) < 1e-2
    # G = [[-cx cy, sx sy], [-sx sy, cx cy]] -> tr(G G) = 2 cx^2 cy^2 - 2 sx^2 sy^2 -> Q = sx^2 sy^2 - cx^2 cy^2
    q = opp.qcrit(u, so)
    xp, yp = so.xp[0][1:-1], so.xp[1][1:-1]
    qa = (np.sin(xp) ** 2)[:, None] * (np.sin(yp) ** 2)[None, :] - (np.cos(xp) ** 2)[:, None] * (np.cos(yp) ** 2)[None, :]
    assert np.max(np.abs(q[1:-1, 1:-1] - qa)) < 5e-3
    # interpolation to pressure points is second order
    up = opp.interpolate_u_p(u, so)
    ua = -np.sin(xp)[:, None] * np.cos(yp)[None, :]
    assert np.max(np.abs(up[1:-1, 1:-1, 0] - ua)) < 1e-3
    wp = opp.interpolate_omega_p(w, so)
    assert np.max(np.abs(wp[1:-1, 1:-1] + 2 * np.sin(xp)[:, None] * np.sin(yp)[None, :])) < 3e-3


def test_3d_vorticity_of_rigid_rotation_and_energy_identity():
    n = 16
    x = tuple(np.linspace(-1.0, 1.0, n + 1) for _ in range(3))
    so = O.Setup(x, {"u": ((O.DirichletBC(), O.DirichletBC()),) * 3})
    u = O.vectorfield(so)
    # u = Omega x r with Omega = (0, 0, 1): (-y, x, 0) sampled at the velocity points -> curl = (0, 0, 2)
    X = [so.xu[a] for a in range(3)]
    u[..., 0] = -X[0][1].reshape(1, -1, 1) + 0 * X[0][0].reshape(-1, 1, 1) + 0 * X[0][2].reshape(1, 1, -1)
    u[..., 1] = X[1][0].reshape(-1, 1, 1) + 0 * X[1][1].reshape(1, -1, 1) + 0 * X[1][2].reshape(1, 1, -1)
    w = opp.vorticity(u, so)
    inner = (slice(1, n), slice(1, n), slice(1, n))
    assert np.allclose(w[inner + (2,)], 2.0, atol=1e-12)
    assert np.allclose(w[inner + (0,)], 0.0, atol=1e-12) and np.allclose(w[inner + (1,)], 0.0, atol=1e-12)
    # ke (not interpolated first) >= ke (interpolated first), equal for a uniform field
    ke_a = opp.kinetic_energy(u, so)
    ke_b = opp.kinetic_energy(u, so, interpolate_first=True)
    assert np.all(ke_a[so.Ip] >= ke_b[so.Ip] - 1e-14)
    u[...] = 1.0
    assert np.allclose(opp.kinetic_energy(u, so)[so.Ip], 1.5) and np.allclose(opp.kinetic_energy(u, so, True)[so.Ip], 1.5)
    # qcrit of rigid rotation: G = [[0,-1,0],[1,0,0],[0,0,0]] -> tr(GG) = -2 -> Q = 1
    u = O.vectorfield(so)
    u[..., 0] = -X[0][1].reshape(1, -1, 1) + 0 * X[0][0].reshape(-1, 1, 1) + 0 * X[0][2].reshape(1, 1, -1)
    u[..., 1] = X[1][0].reshape(-1, 1, 1) + 0 * X[1][1].reshape(1, -1, 1) + 0 * X[1][2].reshape(1, 1, -1)
    q = opp.qcrit(u, so)
    mid = (slice(3, n - 2),) * 3
    assert np.allclose(q[mid], 1.0, atol=1e-12)


def test_energy_spectrum_of_taylor_green_sits_in_one_shell():
    """test/spectral.jl:9-25: TGV energy 1/4 (2-D) and 1/8 (3-D) in the shell kappa = 1."""
    n = 32
    so = _periodic(n, 2)
    u = oini.velocityfield(so, lambda d, x, y: (-np.sin(x) * np.cos(y) if d == 0 else np.cos(x) * np.sin(y)) + 0 * x * y,
                           doproject=False)
    kap, e = opp.energyspectrum(u, so)
    assert abs(e[1] - 0.25) < 1e-12 and abs(e.sum() - 0.25) < 1e-12
    so = _periodic(n, 3)

    def U(d, x, y, z):
        if d == 0:
            return np.sin(x) * np.cos(y) * np.cos(z)
        if d == 1:
            return -np.cos(x) * np.sin(y) * np.cos(z)
        return 0 * x + 0 * y + 0 * z

    u = oini.velocityfield(so, U, doproject=False)
    kap, e = opp.energyspectrum(u, so)
    assert abs(e[1] - 0.125) < 1e-12 and abs(e.sum() - 0.125) < 1e-12
    # Parseval: the shells up to the corner wavenumber hold the mean kinetic energy density
    rng = np.random.default_rng(0)
    u = np.asfortranarray(rng.standard_normal(so.N + (3,)))
    kap, e = opp.energyspectrum(u, so, kmax=int(np.sqrt(3) * n / 2) + 1)
    assert abs(e.sum() - 0.5 * np.mean(np.sum(u[so.Ip] ** 2, axis=-1))) < 1e-12


def test_random_field_spectrum_matches_target_profile():
    """test/spectral.jl:29-56: random_field is divergence free on the staggered grid and its spectrum
    (energyspectrum, an independent restatement) equals the Orlandi target in every shell."""
    from oracle import operators as oops

    for D in (2, 3):
        n = 64 if D == 2 else 32
        x = tuple(np.linspace(0.0, 1.0, n + 1) for _ in range(D))
        so = O.Setup(x, {"u": ((O.PeriodicBC(), O.PeriodicBC()),) * D})
        total, kpeak = 0.3, 5
        u = oini.random_field(so, 0.0, totalenergy=total, kpeak=kpeak, rng=np.random.default_rng(123))
        div = oops.divergence_(O.scalarfield(so), u, so)
        assert np.max(np.abs(div[so.Ip])) < 1e-8
        kdiag = int(np.floor(np.sqrt(sum((m // 2) ** 2 for m in so.Np))))
        kap, e = opp.energyspectrum(u, so, kmax=kdiag)
        target = oini.orlandi_profile(kap.astype(np.float64), kpeak=kpeak)
        target = total * target / target.sum()
        assert abs(e.sum() - total) < 1e-12
        assert np.linalg.norm(e - target) <= 1e-8 * np.linalg.norm(target)


def test_turbulence_statistics_relations_and_taylor_green_closed_form():
    """test/spectral.jl:70-110 (relations between the statistics of a random field, 2-D and 3-D, assertion on
    non-periodic setups) plus a closed form: the 2-D Taylor-Green vortex on [0, 2 pi]^2 has all its energy at
    |k|^2 = 2, so e = 1/4, diss = 2 nu |k|^2 e = nu, enstrophy = 1/2, palinstrophy = |k|^4 e = 1."""
    nu = 1e-3
    for D in (2, 3):
        x = tuple(np.linspace(0.0, 1.0, 33) for _ in range(D))
        so = O.Setup(x, {"u": ((O.PeriodicBC(), O.PeriodicBC()),) * D})
        u = oini.random_field(so, 0.0, totalenergy=0.1, rng=np.random.default_rng(42))
        st = opp.turbulence_statistics(u, so, nu)
        assert all(v > 0 for v in st.values())
        assert abs(st["e"] - 0.1) < 1e-12
        if D == 2:
            assert "enstrophy" in st and "l_kra" in st and "l_kol" not in st
            assert np.isclose(st["uavg"], np.sqrt(st["e"]))
            assert np.isclose(st["diss"], 2 * nu * st["enstrophy"])
            assert np.isclose(st["enstrophy_diss"], 2 * nu * st["palinstrophy"])
        else:
            assert "l_kol" in st and "enstrophy" not in st
            assert np.isclose(st["uavg"], np.sqrt(2 * st["e"] / 3))
            assert np.isclose(st["Re_tay"], st["l_tay"] * st["uavg"] / nu)
            # mean dissipation by Parseval against the spectral derivative of each component
            assert np.isclose(st["kmax_eta"], np.pi * 32 * (nu**3 / st["diss"]) ** 0.25)
    n = 32
    x = tuple(np.linspace(0.0, 2 * np.pi, n + 1) for _ in range(2))
    so = O.Setup(x, {"u": ((O.PeriodicBC(), O.PeriodicBC()),) * 2})
    u = O.vectorfield(so)
    X0, Y0 = np.asarray(so.xu[0][0]).reshape(-1, 1), np.asarray(so.xu[0][1]).reshape(1, -1)
    X1, Y1 = np.asarray(so.xu[1][0]).reshape(-1, 1), np.asarray(so.xu[1][1]).reshape(1, -1)
    u[..., 0] = -np.sin(X0) * np.cos(Y0)
    u[..., 1] = np.cos(X1) * np.sin(Y1)
    st = opp.turbulence_statistics(u, so, nu)
    assert abs(st["e"] - 0.25) < 1e-13 and abs(st["diss"] - nu) < 1e-15
    assert abs(st["enstrophy"] - 0.5) < 1e-12 and abs(st["palinstrophy"] - 1.0) < 1e-12
    assert abs(st["omega_rms"] - 1.0) < 1e-12 and abs(st["Re"] - 0.25 / nu) < 1e-9
    xw = tuple(np.linspace(0.0, 1.0, 17) for _ in range(2))
    sw = O.Setup(xw, {"u": ((O.DirichletBC(), O.DirichletBC()),) * 2})
    with pytest.raises(AssertionError):
        opp.turbulence_statistics(O.vectorfield(sw), sw, nu)
