"""Oracle (test infrastructure): pressure projection and the FFT-type Poisson solvers.

Restates /root/reference/src/pressure.jl:22-78 (poisson!, project, project!,
default_psolver), :215-318 (psolver_transform), :321-382 (psolver_spectral),
:401-545 (psolver_tridiagonal) and the hand-rolled DCT of
/root/reference/src/utils.jl:51-155.

Third-party arithmetic: the reference calls FFTW.jl (compat "1", no Manifest
committed -> version unpinned; Project.toml:44) on the CPU.  FFTW is not in this
image; scipy.fft (pocketfft) provides the same unnormalised forward / 1/n
inverse transforms, float32-preserving.
"""

from __future__ import annotations

import numpy as np
import scipy.fft as sfft

from . import bc as _bc
from . import operators as ops
from .grid import DirichletBC, PeriodicBC, _cospi, _sinpi, scalarfield, vectorfield

WORKERS = None  # scipy.fft workers (set by the CPU-baseline timing leg)


def _fft(a, axis, inverse=False):
    f = sfft.ifft if inverse else sfft.fft
    return f(a, axis=axis, workers=WORKERS)


# ---------------------------------------------------------------------------
# project / project! (pressure.jl:25-55)
# ---------------------------------------------------------------------------
def poisson_(psolver, f):
    """poisson! (pressure.jl:22)."""
    return psolver(f)


def project_(u, setup, psolver, p):
    """project! (pressure.jl:42-55)."""
    T = setup.T
    ops.divergence_(p, u, setup)
    ops.scalewithvolume_(p, setup)
    poisson_(psolver, p)
    _bc.apply_bc_p(p, T(0), setup)
    ops.applypressure_(u, p, setup)
    return u


def project(u, setup, psolver):
    """Allocating project (pressure.jl:25-39)."""
    T = setup.T
    div = ops.divergence_(scalarfield(setup), u, setup)
    div = ops.scalewithvolume_(div, setup)
    p = poisson_(psolver, div.copy(order="F"))
    p = _bc.apply_bc_p(p.copy(order="F"), T(0), setup)
    G = ops.pressuregradient_(vectorfield(setup), p, setup)
    return np.asfortranarray(u - G)


def default_psolver(setup):
    """default_psolver (pressure.jl:58-78).  The direct-solver fallback is out of
    scope: raise instead."""
    D = setup.D
    isperiodic = [setup.is_periodic(a) for a in range(D)]
    iswall = [setup.is_wall(a) for a in range(D)]
    isuniform = [
        bool(np.all(np.isclose(setup.Delta[a][setup.Ip[a]], setup.Delta[a][setup.Ip[a]][0], rtol=np.sqrt(np.finfo(setup.T).eps), atol=0)))
        for a in range(D)
    ]
    ischannel = [
        iswall[a] and all(b == a or (isperiodic[b] and isuniform[b]) for b in range(D))
        for a in range(D)
    ]
    if all(isperiodic) and all(isuniform):
        return psolver_spectral(setup)
    if any(ischannel):
        return psolver_tridiagonal(setup)
    raise NotImplementedError("psolver_direct is outside the hot path (SURVEY.md §2)")


# ---------------------------------------------------------------------------
# psolver_spectral (pressure.jl:321-382)
# ---------------------------------------------------------------------------
def spectral_eigenvalues(setup):
    """ahat tables (pressure.jl:327-342): Float64 evaluation rounded into T."""
    D, T = setup.D, setup.T
    Np = setup.Np
    dx = [setup.Delta[a][0] for a in range(D)]  # first.(Δ): the ghost cell width (:327)
    omega = dx[0]
    for a in range(1, D):
        omega = omega * dx[a]  # prod in T
    kmax = [Np[a] // 2 + 1 if a == 0 else Np[a] for a in range(D)]
    ahat = []
    for a in range(D):
        k = np.arange(kmax[a])
        s2 = _sinpi(k / Np[a]) ** 2  # Float64
        val = (np.float64(T(4) * omega) * s2) / np.float64(dx[a] * dx[a])
        ahat.append(val.astype(T))
    return ahat


def assert_uniform_periodic(setup, name):
    """utils.jl:4-15"""
    assert all(setup.is_periodic(a) for a in range(setup.D)), (
        name + " requires periodic boundary conditions."
    )
    for d in setup.Delta:
        assert np.all(np.abs(d - d[0]) <= max(10 * np.finfo(setup.T).eps, np.sqrt(np.finfo(setup.T).eps) * max(abs(d[0]), np.max(np.abs(d))))), (
            name + " requires uniform grid spacing."
        )
    assert all(n % 2 == 0 for n in setup.N), name + " requires even number of volumes."


def psolver_spectral(setup):
    D, T = setup.D, setup.T
    Ip, Np = setup.Ip, setup.Np
    assert_uniform_periodic(setup, "Spectral psolver")
    ahat = spectral_eigenvalues(setup)
    axes = tuple(range(D - 1, -1, -1))  # real transform along dim 0 (Julia dim 1)
    shape_rev = tuple(Np[a] for a in axes)

    def psolve(p):
        pI = np.ascontiguousarray(p[Ip])  # copyto!(pI, view(p, Ip))
        phat = sfft.rfftn(pI, axes=axes, workers=WORKERS)
        ax = ahat[0].reshape((-1,) + (1,) * (D - 1))
        ay = ahat[1].reshape((1, -1) + (1,) * (D - 2))
        den = ax + ay
        if D == 3:
            den = den + ahat[2].reshape(1, 1, -1)
        with np.errstate(divide="ignore", invalid="ignore"):
            phat = -phat / den
        phat[(0,) * D] = 0
        pI = sfft.irfftn(phat, s=shape_rev, axes=axes, workers=WORKERS)
        p[Ip] = pI.astype(T, copy=False)
        return p

    psolve.kind = "spectral"
    psolve.ahat = ahat
    return psolve


# ---------------------------------------------------------------------------
# manual DCT (utils.jl:51-155)
# ---------------------------------------------------------------------------
def get_perminds(N):
    """utils.jl:51-63 (returned 0-based)."""
    n = N // 2
    assert 2 * n == N, "Only even grids supported"
    x = np.zeros(N, dtype=np.int64)
    xinv = np.zeros(N, dtype=np.int64)
    for k in range(1, n + 1):
        x[k - 1] = 2 * k - 1
        x[N - k] = 2 * k
        xinv[2 * k - 2] = k
        xinv[2 * k - 1] = N - k + 1
    return x - 1, xinv - 1


def manual_dct_stuff(u):
    """utils.jl:65-87: twiddles evaluated in T."""
    T = u.dtype.type
    CT = np.complex64 if T is np.float32 else np.complex128
    w, winv, perm, perminv = [], [], [], []
    for i in range(u.ndim):
        N = u.shape[i]
        k = np.arange(N).astype(T)
        p = T(np.pi)
        arg = (T(2) * p * k / T(4 * N)).astype(T)
        ww = (np.exp(-1j * arg.astype(np.float64)).astype(CT) / np.sqrt(T(2) * T(N)).astype(T)).astype(CT)
        ww[0] = ww[0] / np.sqrt(T(2))
        wi = (np.exp(1j * arg.astype(np.float64)).astype(CT) * np.sqrt(T(2) * T(N)).astype(T)).astype(CT)
        wi[0] = wi[0] * np.sqrt(T(2)) / T(2)
        w.append(ww)
        winv.append(wi)
        pf, pi = get_perminds(N)
        perm.append(pf)
        perminv.append(pi)
    return dict(w=w, winv=winv, perm=perm, perminv=perminv, CT=CT)


def _bshape(ndim, i, n):
    s = [1] * ndim
    s[i] = n
    return s


def manual_dct_(u, i, stuff):
    """manual_dct! (utils.jl:89-122): orthonormal DCT-II along dim i, in place."""
    assert np.isrealobj(u), "Manual DCT only implemented for real inputs"
    T = u.dtype.type
    u[...] = np.take(u, stuff["perm"][i], axis=i)
    uhat = _fft(u.astype(stuff["CT"]), i)
    w = stuff["w"][i].reshape(_bshape(u.ndim, i, u.shape[i]))
    u[...] = (T(2) * np.real(w * uhat)).astype(T)
    return u


def manual_idct_(u, i, stuff):
    """manual_idct! (utils.jl:124-155): orthonormal DCT-III along dim i, in place."""
    assert np.isrealobj(u), "Manual IDCT only implemented for real inputs"
    T = u.dtype.type
    winv = stuff["winv"][i].reshape(_bshape(u.ndim, i, u.shape[i]))
    uhat = (winv * u.astype(stuff["CT"])).astype(stuff["CT"])
    uhat = _fft(uhat, i, inverse=True)
    u[...] = np.real(uhat).astype(T)
    u[...] = np.take(u, stuff["perminv"][i], axis=i)
    return u


# ---------------------------------------------------------------------------
# psolver_transform (pressure.jl:215-318)
# ---------------------------------------------------------------------------
def transform_eigenvalues(setup):
    """pressure.jl:221-246"""
    D, T = setup.D, setup.T
    Np = setup.Np
    dx = [T((setup.xlims[a][1] - setup.xlims[a][0]) / T(Np[a])) for a in range(D)]
    omega = dx[0]
    for a in range(1, D):
        omega = omega * dx[a]
    ahat = []
    for a in range(D):
        n = Np[a]
        k = np.arange(n)
        h2 = np.float64(dx[a] * dx[a])
        if setup.is_periodic(a):
            val = (np.float64(T(-4) * omega) * _sinpi(k / n) ** 2) / h2
        else:
            val = (np.float64(T(2) * omega) * (_cospi(k / n) - 1)) / h2
        ahat.append(val.astype(T))
    return ahat, dx


def psolver_transform(setup):
    D, T = setup.D, setup.T
    Ip, Np = setup.Ip, setup.Np
    for pair in setup.bc_u:
        assert all(isinstance(b, (PeriodicBC, DirichletBC)) for b in pair)
    ahat, dx = transform_eigenvalues(setup)
    for a in range(D):
        d = setup.Delta[a][Ip[a]]
        assert np.all(np.isclose(d, dx[a], rtol=np.sqrt(np.finfo(T).eps), atol=0))
    perdirs = [setup.is_periodic(a) for a in range(D)]
    stuff = manual_dct_stuff(np.zeros(Np, dtype=T))
    CT = stuff["CT"]

    def psolve(pfull):
        p = np.array(pfull[Ip], dtype=T, order="F")
        for i in range(D):
            if not perdirs[i]:
                manual_dct_(p, i, stuff)
        phat = p.astype(CT)
        for i in range(D):
            if perdirs[i]:
                phat = _fft(phat, i)
        ax = ahat[0].reshape((-1,) + (1,) * (D - 1))
        ay = ahat[1].reshape((1, -1) + (1,) * (D - 2))
        den = ax + ay
        if D == 3:
            den = den + ahat[2].reshape(1, 1, -1)
        with np.errstate(divide="ignore", invalid="ignore"):
            phat = (phat / den).astype(CT)
        phat[(0,) * D] = 0
        for i in reversed(range(D)):
            if perdirs[i]:
                phat = _fft(phat, i, inverse=True)
        p = np.asfortranarray(np.real(phat).astype(T))
        for i in reversed(range(D)):
            if not perdirs[i]:
                manual_idct_(p, i, stuff)
        pfull[Ip] = p
        return pfull

    psolve.kind = "transform"
    psolve.ahat = ahat
    return psolve


# ---------------------------------------------------------------------------
# psolver_tridiagonal (pressure.jl:401-545)
# ---------------------------------------------------------------------------
def tridiagonal_tables(setup, dir=None):
    """Coefficient assembly (pressure.jl:406-466)."""
    D, T = setup.D, setup.T
    Ip, Np = setup.Ip, setup.Np
    iswall = [setup.is_wall(a) for a in range(D)]
    isper = [setup.is_periodic(a) for a in range(D)]
    if dir is None:
        if not any(iswall):
            raise ValueError(
                "psolver_tridiagonal: no wall-bounded direction (DirichletBC on both sides) found"
            )
        dir = iswall.index(True)
    if not iswall[dir]:
        raise ValueError(f"psolver_tridiagonal: direction {dir} is not wall-bounded")
    if not all(b == dir or isper[b] for b in range(D)):
        raise ValueError(f"psolver_tridiagonal: all directions except {dir} must be periodic")
    h = [setup.Delta[b][Ip[b].start] for b in range(D)]
    for b in range(D):
        if b == dir:
            continue
        if not np.all(np.isclose(setup.Delta[b][Ip[b]], h[b], rtol=np.sqrt(np.finfo(T).eps), atol=0)):
            raise ValueError(f"psolver_tridiagonal: periodic direction {b} must be uniform")
    n = Np[dir]
    iw = np.arange(Ip[dir].start, Ip[dir].stop)
    a = np.zeros(n, dtype=T)
    b_ = np.zeros(n, dtype=T)
    c = np.zeros(n, dtype=T)
    dl, du = setup.Delta[dir], setup.DeltaU[dir]
    for j in range(n):
        i = iw[j]
        if j != 0:
            a[j] = T(1) / (dl[i] * du[i - 1])
        if j != n - 1:
            c[j] = T(1) / (dl[i] * du[i])
        b_[j] = -(a[j] + c[j])
    lam = []
    for b in range(D):
        if b == dir:
            lam.append(np.zeros(1, dtype=T))
        else:
            k = np.arange(Np[b]).astype(T)
            # sinpi evaluated in T (pressure.jl:457-458)
            sp = _sinpi((k / T(Np[b])).astype(np.float64)).astype(T)
            lam.append((T(-4) * sp**2 / h[b] ** 2).astype(T))
    hper = T(1)
    for b in range(D):
        hper = hper * (T(1) if b == dir else h[b])
    invO = (T(1) / (hper * dl[iw])).astype(T)
    return dict(dir=dir, a=a, b=b_, c=c, lam=lam, invO=invO, n=n)


def psolver_tridiagonal(setup, dir=None):
    D, T = setup.D, setup.T
    Ip, Np = setup.Ip, setup.Np
    tb = tridiagonal_tables(setup, dir)
    dir = tb["dir"]
    a, b, c, lam, n = tb["a"], tb["b"], tb["c"], tb["lam"], tb["n"]
    CT = np.complex64 if T is np.float32 else np.complex128
    invO = tb["invO"].reshape(_bshape(D, dir, n))
    fftscale = T(np.prod([1 if be == dir else Np[be] for be in range(D)]))

    def line(j):
        return tuple(j if be == dir else slice(None) for be in range(D))

    def psolve(p):
        pI = np.array(p[Ip], dtype=T, order="F")
        phat = (pI * invO).astype(CT)
        for be in range(D):
            if be != dir:
                phat = _fft(phat, be)
        # Thomas kernel (pressure.jl:480-504), vectorised over the modes
        lamsum = np.zeros([1 if be == dir else Np[be] for be in range(D)], dtype=T)
        for be in range(D):
            if be != dir:
                lamsum = lamsum + lam[be].reshape(_bshape(D, be, Np[be]))
        lam2 = lamsum[line(0)]  # drop dir axis
        w = np.zeros(Np, dtype=T, order="F")
        zero = lam2 == 0
        with np.errstate(divide="ignore", invalid="ignore"):
            w[line(0)] = np.where(zero, T(0), c[0] / (b[0] + lam2))
            phat[line(0)] = np.where(zero, CT(0), phat[line(0)] / (b[0] + lam2))
        for j in range(1, n):
            denom = b[j] + lam2 - a[j] * w[line(j - 1)]
            w[line(j)] = c[j] / denom
            phat[line(j)] = (phat[line(j)] - a[j] * phat[line(j - 1)]) / denom
        for j in range(n - 2, -1, -1):
            phat[line(j)] -= w[line(j)] * phat[line(j + 1)]
        # zero-mode mean removal (:531-532)
        m0 = tuple(slice(None) if be == dir else 0 for be in range(D))
        phat[m0] -= np.sum(phat[m0]) / n
        # unnormalised backward FFT (:535-538)
        for be in range(D):
            if be != dir:
                phat = _fft(phat, be, inverse=True) * CT(Np[be])
        pI = (np.real(phat) / fftscale).astype(T)
        p[Ip] = pI
        return p

    psolve.kind = "tridiagonal"
    psolve.fftscale = fftscale
    psolve.tables = tb
    return psolve


# ---------------------------------------------------------------------------
# psolver_cg (pressure.jl:135-203)
# ---------------------------------------------------------------------------
def create_laplace_diag(setup):
    """create_laplace_diag (pressure.jl:136-153): z[I] = -p[I] / d over Ip."""
    D, T = setup.D, setup.T
    Ip = setup.Ip

    def met(arr, dim, off=0):
        v = arr[Ip[dim].start + off : Ip[dim].stop + off]
        shape = [1] * D
        shape[dim] = len(v)
        return v.reshape(shape)

    omega = met(setup.Delta[0], 0)
    for a in range(1, D):
        omega = omega * met(setup.Delta[a], a)
    d = np.zeros(setup.Np, dtype=T)
    for a in range(D):
        d = d - omega / met(setup.Delta[a], a) * (T(1) / met(setup.DeltaU[a], a) + T(1) / met(setup.DeltaU[a], a, -1))

    def laplace_diag(z, p):
        z[Ip] = -p[Ip] / d
        return z

    return laplace_diag


def psolver_cg(setup, abstol=0.0, reltol=None, maxiter=None, preconditioner=None):
    """psolver_cg (pressure.jl:156-203): diagonally preconditioned conjugate gradients."""
    T = setup.T
    Ip = setup.Ip
    reltol = np.sqrt(np.finfo(T).eps) if reltol is None else reltol
    maxiter = int(np.prod(setup.Np)) if maxiter is None else maxiter
    preconditioner = preconditioner or create_laplace_diag(setup)
    r, L, q = scalarfield(setup), scalarfield(setup), scalarfield(setup)

    def psolve(p):
        q[...] = 0
        r[...] = p
        rho_prev = T(1)
        residual = np.sqrt(np.sum(r[Ip] ** 2))
        tolerance = max(T(reltol) * residual, T(abstol))
        iteration = 0
        p[...] = 0
        while iteration < maxiter and residual > tolerance:
            preconditioner(L, r)
            rho = np.sum(L[Ip] * r[Ip])
            beta = rho / rho_prev
            q[...] = L + beta * q
            _bc.apply_bc_p(q, T(0), setup)
            ops.laplacian_(L, q, setup)
            alpha = rho / np.sum(q[Ip] * L[Ip])
            p += alpha * q
            r[...] = r - alpha * L
            rho_prev = rho
            residual = np.sqrt(np.sum(r[Ip] ** 2))
            iteration += 1
        psolve.iterations = iteration
        return p

    psolve.kind = "cg"
    return psolve
