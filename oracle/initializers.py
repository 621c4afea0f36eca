"""Oracle (test infrastructure): input-field generators.

Restates /root/reference/src/initializers.jl:45-87 (velocityfield,
temperaturefield) and :109-183 (random_field) with the shell bookkeeping of
/root/reference/src/spectral.jl:7-67.  Not timed; needed so that the oracle and
the CUDA path are fed identical fields.  Julia's Xoshiro/randn streams cannot be
reproduced outside Julia; NumPy ``default_rng`` is used instead and the SAME
arrays go to both sides.
"""

from __future__ import annotations

import numpy as np
import scipy.fft as sfft

from . import bc as _bc
from .grid import scalarfield, vectorfield
from .pressure import assert_uniform_periodic, default_psolver, project


def _bcoords(D, xs, slices):
    out = []
    for b in range(D):
        c = xs[b][slices[b]]
        shape = [1] * D
        shape[b] = len(c)
        out.append(c.reshape(shape))
    return out


def velocityfield(setup, ufunc, t=0.0, psolver=None, doproject=True):
    """initializers.jl:45-77.  ``ufunc(dim, x...)`` with a 0-based ``dim``."""
    D, T = setup.D, setup.T
    u = vectorfield(setup)
    for a in range(D):
        xin = _bcoords(D, setup.xu[a], setup.Iu[a])
        u[tuple(setup.Iu[a]) + (a,)] = np.asarray(ufunc(a, *xin), dtype=T)
    _bc.apply_bc_u(u, T(t), setup)
    if doproject:
        psolver = psolver if psolver is not None else default_psolver(setup)
        u = project(u, setup, psolver)
        _bc.apply_bc_u(u, T(t), setup)
    return u


def temperaturefield(setup, tempfunc, t=0.0):
    """initializers.jl:80-87"""
    D, T = setup.D, setup.T
    xin = _bcoords(D, setup.xp, setup.Ip)
    temp = scalarfield(setup)
    temp[setup.Ip] = np.asarray(tempfunc(*xin), dtype=T)
    return _bc.apply_bc_temp(temp, T(t), setup)


def orlandi_profile(k, kpeak=10):
    """initializers.jl:93"""
    return k**4 * np.exp(-2 * (k / kpeak) ** 2)


def fftfreq_int(n, i):
    """spectral.jl:7 (i is 1-based)."""
    return i - 1 - (0 if i <= (n + 1) >> 1 else n)


def random_field(setup, t=0.0, profile=orlandi_profile, totalenergy=1.0, rng=None, **kwargs):
    """initializers.jl:109-183"""
    D, T = setup.D, setup.T
    Np, Ip = setup.Np, setup.Ip
    rng = np.random.default_rng() if rng is None else rng
    CT = np.complex64 if T is np.float32 else np.complex128
    assert_uniform_periodic(setup, "Random field")
    axes = tuple(range(D - 1, -1, -1))
    uhat = [
        sfft.rfftn(rng.standard_normal(Np).astype(T), axes=axes).astype(CT) for _ in range(D)
    ]
    # staggered-divergence symbol (:133-139)
    m = []
    kints = []
    for a in range(D):
        L = setup.xlims[a][1] - setup.xlims[a][0]
        h = T(L / Np[a])
        if a == 0:
            kint = np.arange(Np[0] // 2 + 1)
        else:
            kint = np.array([fftfreq_int(Np[a], i) for i in range(1, Np[a] + 1)])
        kints.append(kint)
        ma = ((1 - np.exp(-1j * (T(2 * np.pi) * kint.astype(T) / T(Np[a])).astype(np.float64))) / h).astype(CT)
        shape = [1] * D
        shape[a] = len(kint)
        m.append(ma.reshape(shape))
    mdotu = m[0] * uhat[0]
    for a in range(1, D):
        mdotu = mdotu + m[a] * uhat[a]
    m2 = sum(np.abs(ma) ** 2 for ma in m)
    with np.errstate(divide="ignore", invalid="ignore"):
        for a in range(D):
            uhat[a] = uhat[a] - np.where(m2 > 0, np.conj(m[a]) * mdotu / m2, 0)
    efield = sum(np.abs(uh) ** 2 for uh in uhat)
    # shells (spectral.jl:39-67)
    kdiag = int(np.floor(np.sqrt(sum((Np[a] // 2) ** 2 for a in range(D)))))
    k2 = None
    for a in range(D):
        shape = [1] * D
        shape[a] = len(kints[a])
        term = (kints[a] ** 2).reshape(shape)
        k2 = term if k2 is None else k2 + term
    shell = np.floor(np.sqrt(k2)).astype(np.int64)
    Kx = Np[0] // 2 + 1
    ix = np.arange(Kx).reshape((-1,) + (1,) * (D - 1))
    weight = np.where((ix == 0) | (ix == Kx - 1), 1.0, 2.0)  # conjugate modes counted twice
    totalprofile = sum(T(profile(T(k), **kwargs)) for k in range(kdiag + 1))
    Ntot = T(np.prod(Np))
    for kappa in range(kdiag + 1):
        mask = shell == kappa
        eshell = np.sum(efield[mask] * np.broadcast_to(weight, efield.shape)[mask]) / (2 * Ntot**2)
        e0 = T(totalenergy) * T(profile(T(kappa), **kwargs)) / totalprofile
        factor = np.sqrt(e0 / eshell) if eshell > 0 else T(0)
        for a in range(D):
            uhat[a][mask] *= factor
    u = vectorfield(setup)
    shape_rev = tuple(Np[a] for a in axes)
    for a in range(D):
        v = sfft.irfftn(uhat[a], s=shape_rev, axes=axes)
        u[tuple(Ip) + (a,)] = v.astype(T)
    _bc.apply_bc_u(u, T(t), setup)
    return u
