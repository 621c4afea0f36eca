"""Oracle pinned against /root/reference/test/psolvers.jl and test/dct.jl. CPU only."""
import numpy as np
import pytest
import scipy.fft as sfft

import oracle as O
from oracle import bc, operators as ops, pressure

from oracle_helpers import periodic_setup


def _lap_of(setup, p_exact):
    bc.apply_bc_p(p_exact, 0.0, setup)
    return ops.laplacian_(O.scalarfield(setup), p_exact, setup)


@pytest.mark.parametrize("T", [np.float64, np.float32])
def test_manufactured_pressure_spectral_and_transform(T):
    # test/psolvers.jl:1-37: p = 1/4 (cos 2x + cos 2y) on 32^2 periodic [0, 2pi]^2
    n = 32
    setup = periodic_setup(n, 2, T, L=2 * np.pi)
    xp = setup.xp
    p_exact = np.asfortranarray(
        (0.25 * (np.cos(2 * xp[0])[:, None] + np.cos(2 * xp[1])[None, :])).astype(T)
    )
    lap = _lap_of(setup, p_exact)
    rtol = np.sqrt(np.finfo(T).eps) if T is np.float64 else 2e-3
    for mk in (pressure.psolver_spectral, pressure.psolver_transform):
        p = bc.apply_bc_p(pressure.poisson_(mk(setup), lap.copy(order="F")), 0.0, setup)
        err = np.linalg.norm(p[setup.Ip] - p_exact[setup.Ip]) / np.linalg.norm(p_exact[setup.Ip])
        assert err < rtol, (mk.__name__, err)
    assert pressure.default_psolver(setup).kind == "spectral"


@pytest.mark.parametrize("D", [2, 3])
def test_tridiagonal_channel(D):
    # test/psolvers.jl:39-83: walls in dir (cosine grid), periodic uniform elsewhere
    for dir in range(D):
        n = 16
        x = tuple(
            O.cosine_grid(0.0, 1.0, n) if b == dir else np.linspace(0, 2 * np.pi, 2 * n + 1)
            for b in range(D)
        )
        bcs = dict(
            u=tuple(
                (O.DirichletBC(), O.DirichletBC()) if b == dir else (O.PeriodicBC(), O.PeriodicBC())
                for b in range(D)
            )
        )
        setup = O.Setup(x, bcs)
        rng = np.random.default_rng(129 + D + dir)
        p_exact = O.scalarfield(setup)
        p_exact[setup.Ip] = rng.standard_normal(setup.Np)
        lap = _lap_of(setup, p_exact)
        tri = pressure.psolver_tridiagonal(setup)
        p = pressure.poisson_(tri, lap.copy(order="F"))
        a = p[setup.Ip] - p[setup.Ip].mean()
        b = p_exact[setup.Ip] - p_exact[setup.Ip].mean()
        assert np.allclose(a, b, rtol=1e-8, atol=1e-10)
        explicit = pressure.psolver_tridiagonal(setup, dir=dir)
        assert np.allclose(explicit(lap.copy(order="F"))[setup.Ip], p[setup.Ip])
        with pytest.raises(ValueError):
            pressure.psolver_tridiagonal(setup, dir=(dir + 1) % D)
        assert pressure.default_psolver(setup).kind == "tridiagonal"
    with pytest.raises(ValueError):
        pressure.psolver_tridiagonal(periodic_setup(16, 2))


def test_transform_with_walls_3d():
    # psolver_transform on periodic-x / wall-y / wall-z (BASELINE config 4 shape)
    n = (16, 12, 10)
    x = tuple(np.linspace(0, L, m + 1) for L, m in zip((np.pi, 1.0, 1.0), n))
    P, W = (O.PeriodicBC(), O.PeriodicBC()), (O.DirichletBC(), O.DirichletBC())
    setup = O.Setup(x, dict(u=(P, W, W)))
    rng = np.random.default_rng(5)
    p_exact = O.scalarfield(setup)
    p_exact[setup.Ip] = rng.standard_normal(setup.Np)
    p_exact[setup.Ip] -= p_exact[setup.Ip].mean()
    lap = _lap_of(setup, p_exact)
    p = pressure.psolver_transform(setup)(lap.copy(order="F"))
    assert np.allclose(p[setup.Ip], p_exact[setup.Ip], rtol=1e-9, atol=1e-11)


def test_perminds():
    # test/dct.jl:1-12
    perm, perminv = pressure.get_perminds(10)
    u = np.random.default_rng(0).standard_normal(10)
    assert np.array_equal(u[perm][perminv], u) and np.array_equal(u[perminv][perm], u)
    with pytest.raises(AssertionError):
        pressure.get_perminds(11)


@pytest.mark.parametrize("shape", [(30,), (20, 26), (10, 16, 28)])
def test_manual_dct_matches_library(shape):
    # test/dct.jl:14-44: manual_dct!/idct! == FFTW dct/idct (orthonormal DCT-II/III)
    u = np.asfortranarray(np.random.default_rng(1).standard_normal(shape))
    stuff = pressure.manual_dct_stuff(u)
    for i in range(len(shape)):
        assert np.allclose(
            pressure.manual_dct_(u.copy(order="F"), i, stuff), sfft.dct(u, type=2, norm="ortho", axis=i)
        )
        assert np.allclose(
            pressure.manual_idct_(u.copy(order="F"), i, stuff), sfft.idct(u, type=2, norm="ortho", axis=i)
        )
    with pytest.raises(AssertionError):
        pressure.manual_dct_(u + 1j, 0, stuff)


def test_psolver_cg_reproduces_manufactured_pressure():
    """test/psolvers.jl:13-36 for psolver_cg, on grids / BCs the FFT-type solvers do not cover."""
    import oracle as O
    from oracle import bc as obc, operators as oops, pressure as opr

    cases = [
        ((O.tanh_grid(0, 1, 12, 1.2), O.tanh_grid(0, 1, 10, 1.1), O.cosine_grid(0, 1, 14)),
         ((O.DirichletBC(), O.DirichletBC()),) * 3, False),
        ((np.linspace(0, 1, 11), O.cosine_grid(0, 1, 8), np.linspace(0, 1, 13)),
         ((O.DirichletBC(), O.PressureBC()), (O.SymmetricBC(), O.SymmetricBC()), (O.PeriodicBC(), O.PeriodicBC())), True),
    ]
    for x, bc, pinned in cases:
        so = O.Setup(tuple(np.asarray(v) for v in x), {"u": bc})
        pe = O.scalarfield(so)
        pe[so.Ip] = np.random.default_rng(3).standard_normal(so.Np)
        obc.apply_bc_p(pe, 0.0, so)
        lap = oops.laplacian_(O.scalarfield(so), pe, so)
        ps = opr.psolver_cg(so)
        got = ps(lap.copy(order="F"))[so.Ip]
        want = pe[so.Ip]
        if not pinned:
            got, want = got - got.mean(), want - want.mean()
        assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-5
        assert 0 < ps.iterations < np.prod(so.Np)
