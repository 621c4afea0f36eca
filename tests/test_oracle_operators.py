"""Oracle pinned against the reference's own operator tests
(/root/reference/test/operators.jl).  CPU only."""
import numpy as np
import pytest

import oracle as O
from oracle import operators as ops
from oracle import bc, pressure, initializers


def _setup(D, T=np.float64):
    # test/operators.jl:1-45: n = 16, non-uniform grids, all-Dirichlet
    n = 16
    if D == 2:
        x = (O.tanh_grid(0, 1, n, dtype=T), O.tanh_grid(0, 1, n, 1.3, dtype=T))
    else:
        x = (
            O.tanh_grid(0, 1, n, 1.2, dtype=T),
            O.tanh_grid(0, 1, n, 1.1, dtype=T),
            O.cosine_grid(0, 1, n, dtype=T),
        )
    b = (O.DirichletBC(), O.DirichletBC())
    setup = O.Setup(x, dict(u=(b,) * D, temp=(b,) * D))
    return setup


def _uref(dim, x, y, *a):
    return -(dim == 0) * np.sin(x) * np.cos(y) + (dim == 1) * np.cos(x) * np.sin(y) + 0 * sum(a)


def _psolver_cg(setup):
    """Test-only stand-in for the reference's psolver_direct (out of scope for
    the product): matrix-free CG on -L, L = laplacian! (operators.jl:320-359)."""
    from scipy.sparse.linalg import LinearOperator, cg

    Ip = setup.Ip
    shape = tuple(s.stop - s.start for s in Ip)
    n = int(np.prod(shape))

    def mv(x):
        p = O.scalarfield(setup)
        p[Ip] = x.reshape(shape, order="F")
        bc.apply_bc_p(p, 0.0, setup)
        return -ops.laplacian_(O.scalarfield(setup), p, setup)[Ip].reshape(-1, order="F")

    A = LinearOperator((n, n), matvec=mv, dtype=np.float64)

    def psolve(p):
        f = -p[Ip].reshape(-1, order="F")
        f = f - f.mean()
        x, info = cg(A, f, rtol=1e-15, atol=0, maxiter=20000)
        p[Ip] = (x - x.mean()).reshape(shape, order="F")
        return p

    return psolve


_CACHE = {}


def _field(setup, rng):
    # test/operators.jl:23,43: u = velocityfield(setup, uref, 0) (projected)
    key = setup.D
    if key not in _CACHE:
        _CACHE[key] = initializers.velocityfield(setup, _uref, psolver=_psolver_cg(setup))
    return _CACHE[key]


def _w(setup, comp):
    """Volume weights of the velocity DOFs of component comp (test/operators.jl:73-81)."""
    D = setup.D
    w = 1
    for b in range(D):
        arr = setup.DeltaU[b] if b == comp else setup.Delta[b]
        shape = [1] * D
        shape[b] = -1
        w = w * arr.reshape(shape)
    return w


@pytest.mark.parametrize("D", [2, 3])
def test_divergence_is_minus_gradient_transpose(D):
    # test/operators.jl:54-84
    setup = _setup(D)
    rng = np.random.default_rng(0)
    v = np.asfortranarray(rng.standard_normal(setup.N + (D,)))
    p = np.asfortranarray(rng.standard_normal(setup.N))
    bc.apply_bc_u(v, 0.0, setup)
    bc.apply_bc_p(p, 0.0, setup)
    Dv = ops.divergence_(O.scalarfield(setup), v, setup)
    Gp = ops.pressuregradient_(O.vectorfield(setup), p, setup)
    ODv = ops.scalewithvolume_(Dv.copy(order="F"), setup)
    pDv = np.sum((p * ODv)[setup.Ip])
    vGp = 0.0
    for a in range(D):
        vGp += np.sum((v[..., a] * _w(setup, a) * Gp[..., a])[setup.Iu[a]])
    assert np.isclose(pDv, -vGp, rtol=1e-10)


@pytest.mark.parametrize("D", [2, 3])
def test_convection_skew_symmetric(D):
    # test/operators.jl:103-124
    setup = _setup(D)
    u = _field(setup, None)
    c = ops.convection_(np.zeros_like(u), u, setup)
    uCu = 0.0
    for a in range(D):
        uCu += np.sum((u[..., a] * _w(setup, a) * c[..., a])[setup.Iu[a]])
    assert abs(uCu) < 1e-12


@pytest.mark.parametrize("D", [2, 3])
def test_diffusion_dissipative(D):
    # test/operators.jl:126-150
    setup = _setup(D)
    u = _field(setup, None)
    d = ops.diffusion_(np.zeros_like(u), u, setup, 1e-3)
    uDu = 0.0
    for a in range(D):
        uDu += np.sum((u[..., a] * _w(setup, a) * d[..., a])[setup.Iu[a]])
    assert uDu <= 0


@pytest.mark.parametrize("D", [2, 3])
def test_convectiondiffusion_equals_sum(D):
    # test/operators.jl:152-167
    setup = _setup(D)
    u = _field(setup, None)
    cd = ops.convectiondiffusion_(np.zeros_like(u), u, setup, 1e-3)
    c = ops.convection_(np.zeros_like(u), u, setup)
    d = ops.diffusion_(np.zeros_like(u), u, setup, 1e-3)
    assert np.allclose(cd, c + d, rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("D", [2, 3])
def test_laplacian_negative_and_symmetric(D):
    # test/operators.jl:86-101 (negativity) + L = Ω D G identity used by the psolvers
    setup = _setup(D)
    rng = np.random.default_rng(1)
    p = np.asfortranarray(rng.standard_normal(setup.N))
    bc.apply_bc_p(p, 0.0, setup)
    Lp = ops.laplacian_(O.scalarfield(setup), p, setup)
    assert np.sum((p * Lp)[setup.Ip]) <= 0
    # laplacian! already carries the volume factor: Lp == Ω div(grad p) on Ip
    # when the wall-normal gradient at the walls is dropped (homogeneous Neumann).
    G = ops.pressuregradient_(O.vectorfield(setup), p, setup)
    bc.apply_bc_u(G, 0.0, setup)  # no-slip zeroes the wall-normal faces
    DG = ops.scalewithvolume_(ops.divergence_(O.scalarfield(setup), G, setup), setup)
    assert np.allclose(Lp[setup.Ip], DG[setup.Ip], rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("D", [2, 3])
def test_gravity_only_in_gdir_and_rhs_finite(D):
    # test/operators.jl:169-205
    setup = _setup(D)
    rng = np.random.default_rng(2)
    u = _field(setup, None)
    temp = np.asfortranarray(rng.standard_normal(setup.N))
    g = ops.applygravity_(np.zeros_like(u), temp, setup, 1, 1.0)
    for a in range(D):
        if a != 1:
            assert np.all(g[..., a] == 0)
    assert np.any(g[..., 1] != 0)
    f = dict(u=np.zeros_like(u), temp=np.zeros_like(temp))
    ops.boussinesq_(f, dict(u=u, temp=temp), 0.0, setup, 1e-3, 1e-3, 1, 1.0, True)
    assert np.all(np.isfinite(f["u"])) and np.all(np.isfinite(f["temp"]))


def test_grids():
    # test/grid.jl: endpoints and monotonicity
    for g in (O.cosine_grid(0.0, 2.0, 10), O.tanh_grid(0.0, 2.0, 10, 1.4), O.stretched_grid(0.0, 2.0, 10, 1.1)):
        assert np.isclose(g[0], 0) and np.isclose(g[-1], 2)
        assert np.all(np.diff(g) > 0)


def _linear_u(setup, A, c):
    """u_a = c_a + sum_b A[a, b] x_b at the velocity points of component a."""
    D = setup.D
    u = O.vectorfield(setup)
    for a in range(D):
        acc = c[a]
        for b in range(D):
            shape = [1] * D
            shape[b] = -1
            acc = acc + A[a, b] * np.asarray(setup.xu[a][b], dtype=np.float64).reshape(shape)
        u[..., a] = acc
    return u


def _scalar_at_p(setup, fun):
    D = setup.D
    xs = []
    for b in range(D):
        shape = [1] * D
        shape[b] = -1
        xs.append(np.asarray(setup.xp[b], dtype=np.float64).reshape(shape))
    out = O.scalarfield(setup)
    out[...] = fun(*xs)
    return out


@pytest.mark.parametrize("D", [2, 3])
def test_boussinesq_terms_closed_forms_on_stretched_grids(D):
    """Known answers for the three Boussinesq kernels on the reference's stretched test grids (the reference only
    checks that they run, test/operators.jl:169-205): for a linear velocity field the collocated gradient is exact, so
    dissipation! gives nu |grad u|_F^2; for a linear temperature avg() is exact at the faces, so applygravity! gives
    g T(x_face) and convection_diffusion_temp! gives -div(u T) exactly when u is constant along its own direction
    (zero diffusion of a linear field)."""
    setup = _setup(D)
    rng = np.random.default_rng(11)
    inner = (slice(2, -2),) * D  # the collocation reaches one cell to the left
    # dissipation!
    A = rng.standard_normal((D, D))
    u = _linear_u(setup, A, rng.standard_normal(D))
    diss = ops.dissipation_(O.scalarfield(setup), u, setup, 0.37)
    np.testing.assert_allclose(diss[inner], 0.37 * np.sum(A * A), rtol=1e-11)
    # applygravity!
    g = rng.standard_normal(D)
    temp = _scalar_at_p(setup, lambda *x: 0.3 + sum(gi * xi for gi, xi in zip(g, x)))
    for gdir in range(D):
        f = ops.applygravity_(O.vectorfield(setup), temp, setup, gdir, 1.7)
        xs = []
        for b in range(D):
            shape = [1] * D
            shape[b] = -1
            xs.append(np.asarray(setup.xu[gdir][b], dtype=np.float64).reshape(shape))
        want = 1.7 * (0.3 + sum(gi * xi for gi, xi in zip(g, xs)))
        sl = (slice(1, -1),) * D
        np.testing.assert_allclose(f[sl + (gdir,)], np.broadcast_to(want, f[..., gdir].shape)[sl], rtol=1e-12)
        for a in range(D):
            if a != gdir:
                assert np.all(f[..., a] == 0)
    # convection_diffusion_temp!: u_a = c_a + sum_{b != a} A[a, b] x_b (divergence free, constant along a)
    A0 = A - np.diag(np.diag(A))
    c0 = rng.standard_normal(D)
    u = _linear_u(setup, A0, c0)
    ct = ops.convection_diffusion_temp_(O.scalarfield(setup), u, temp, setup, 0.21)
    xs = []
    for b in range(D):
        shape = [1] * D
        shape[b] = -1
        xs.append(np.asarray(setup.xp[b], dtype=np.float64).reshape(shape))
    want = 0
    for a in range(D):  # -d/dx_a (u_a T) = -u_a g_a with u_a at the cell centre (it does not depend on x_a)
        ua = c0[a] + sum(A0[a, b] * xs[b] for b in range(D))
        want = want - ua * g[a]
    sl = (slice(1, -1),) * D
    np.testing.assert_allclose(ct[sl], np.broadcast_to(want, ct.shape)[sl], rtol=1e-10, atol=1e-12)


def test_temperature_diffusion_of_a_parabola_is_exact_on_a_uniform_grid():
    """kappa * laplace(T) for T = sum q_a x_a^2 is 2 kappa sum(q) exactly (central differences of a quadratic), u = 0."""
    x = tuple(np.linspace(0.0, L, n + 1) for n, L in ((8, 1.0), (10, 2.0), (6, 0.5)))
    b = (O.PeriodicBC(), O.PeriodicBC())
    setup = O.Setup(x, dict(u=(b,) * 3, temp=(b,) * 3))
    q = np.array([0.7, -1.3, 2.1])
    temp = _scalar_at_p(setup, lambda X, Y, Z: q[0] * X**2 + q[1] * Y**2 + q[2] * Z**2)
    ct = ops.convection_diffusion_temp_(O.scalarfield(setup), O.vectorfield(setup), temp, setup, 0.4)
    inner = (slice(2, -2),) * 3  # away from the periodic seam, where the parabola is not periodic
    np.testing.assert_allclose(ct[inner], 0.4 * 2 * q.sum(), rtol=1e-11)


def test_momentum_operators_closed_forms_on_a_uniform_grid():
    """convection! of a divergence-free linear field u = c + A x is -(u . grad) u = -A (c + A x) exactly (the face
    interpolations are exact for linear fields, the flux differences are central differences of quadratics);
    diffusion! of a quadratic field is nu * 2 sum(q) exactly; convectiondiffusion! is their sum up to rounding."""
    rng = np.random.default_rng(5)
    x = tuple(np.linspace(0.0, L, n + 1) for n, L in ((10, 1.0), (8, 2.0), (12, 1.5)))
    b = (O.PeriodicBC(), O.PeriodicBC())
    setup = O.Setup(x, dict(u=(b,) * 3))
    A = rng.standard_normal((3, 3))
    A -= np.trace(A) / 3 * np.eye(3)
    c = rng.standard_normal(3)
    u = _linear_u(setup, A, c)
    conv = ops.convection_(O.vectorfield(setup), u, setup)
    inner = (slice(2, -2),) * 3  # away from the periodic seam (a linear field is not periodic)
    AA = A @ A
    for i in range(3):
        xs = []
        for d in range(3):
            shape = [1, 1, 1]
            shape[d] = -1
            xs.append(np.asarray(setup.xu[i][d], dtype=np.float64).reshape(shape))
        want = -(A[i] @ c) - sum(AA[i, d] * xs[d] for d in range(3))
        np.testing.assert_allclose(conv[inner + (i,)], np.broadcast_to(want, conv[..., i].shape)[inner], rtol=1e-10, atol=1e-11)
    q = rng.standard_normal((3, 3))
    uq = O.vectorfield(setup)
    for i in range(3):
        acc = 0
        for d in range(3):
            shape = [1, 1, 1]
            shape[d] = -1
            acc = acc + q[i, d] * np.asarray(setup.xu[i][d], dtype=np.float64).reshape(shape) ** 2
        uq[..., i] = acc
    diff = ops.diffusion_(O.vectorfield(setup), uq, setup, 0.3)
    for i in range(3):
        np.testing.assert_allclose(diff[inner + (i,)], 0.3 * 2 * q[i].sum(), rtol=1e-10)
    both = ops.convectiondiffusion_(O.vectorfield(setup), u, setup, 0.3)  # diffusion of a linear field is zero
    np.testing.assert_allclose(both[inner], conv[inner], rtol=1e-10, atol=1e-11)


def test_second_order_accuracy_of_the_momentum_right_hand_side():
    """convectiondiffusion! against the analytic -div(u u) + nu lap(u) of the Taylor-Green field
    u = (sin x cos y, -cos x sin y): -(u . grad) u = -(sin 2x, sin 2y) / 2, lap(u) = -2 u.  The error falls by 4 per
    halving of h (second-order scheme), and the divergence / pressure gradient / Laplacian chain likewise."""
    nu = 0.1
    errs, lerrs = [], []
    for n in (16, 32, 64):
        x = tuple(np.linspace(0.0, 2 * np.pi, n + 1) for _ in range(2))
        b = (O.PeriodicBC(), O.PeriodicBC())
        setup = O.Setup(x, dict(u=(b, b)))
        u = O.vectorfield(setup)
        X0, Y0 = np.asarray(setup.xu[0][0]).reshape(-1, 1), np.asarray(setup.xu[0][1]).reshape(1, -1)
        X1, Y1 = np.asarray(setup.xu[1][0]).reshape(-1, 1), np.asarray(setup.xu[1][1]).reshape(1, -1)
        u[..., 0] = np.sin(X0) * np.cos(Y0)
        u[..., 1] = -np.cos(X1) * np.sin(Y1)
        f = ops.convectiondiffusion_(O.vectorfield(setup), u, setup, nu)
        want0 = -0.5 * np.sin(2 * X0) - 2 * nu * np.sin(X0) * np.cos(Y0) + 0 * Y0
        want1 = -0.5 * np.sin(2 * Y1) + 2 * nu * np.cos(X1) * np.sin(Y1) + 0 * X1
        inner = (slice(1, -1),) * 2
        errs.append(max(np.abs(f[inner + (0,)] - want0[inner]).max(), np.abs(f[inner + (1,)] - want1[inner]).max()))
        # laplacian! = Omega div grad: p = cos x cos y -> lap p = -2 p (times the cell volume)
        p = O.scalarfield(setup)
        XP, YP = np.asarray(setup.xp[0]).reshape(-1, 1), np.asarray(setup.xp[1]).reshape(1, -1)
        p[...] = np.cos(XP) * np.cos(YP)
        Lp = ops.laplacian_(O.scalarfield(setup), p, setup)
        h = 2 * np.pi / n
        lerrs.append(np.abs(Lp[inner] / h**2 + 2 * p[inner]).max())
    for e in (errs, lerrs):
        rates = [np.log2(e[i] / e[i + 1]) for i in range(2)]
        assert 1.9 < min(rates) and max(rates) < 2.1, (e, rates)
