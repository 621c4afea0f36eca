"""Oracle (test infrastructure): field diagnostics of /root/reference/src/operators.jl
(vorticity!, interpolate_u_p!, interpolate_ω_p!, qcrit!, kinetic_energy!, total_kinetic_energy),
restated with NumPy slicing in the reference's evaluation order."""
from __future__ import annotations

import numpy as np

from . import scalarfield, vectorfield
from .operators import _e, _m, _sl, _sub, _zero, delta_u, scalewithvolume_


def _rng(setup, lo, hi, off, comp=None):
    """slices lo[a]+off[a] : hi[a]+off[a] (0-based, half open) [+ component]"""
    s = tuple(slice(l + o, h + o) for l, h, o in zip(lo, hi, off))
    return s if comp is None else s + (comp,)


def _tab(setup, arr, dim, lo, hi):
    v = arr[lo:hi]
    shape = [1] * setup.D
    shape[dim] = len(v)
    return v.reshape(shape)


def vorticity_(omega, u, setup):
    """vorticity! (operators.jl:853-881): ndrange = N .- 1, no offset."""
    D, N = setup.D, setup.N
    lo, hi = (0,) * D, tuple(n - 1 for n in N)
    z = _zero(D)

    def d(comp, dim):  # (u[right(I, dim), comp] - u[I, comp]) / Δu[dim][I[dim]]
        return (u[_rng(setup, lo, hi, _e(D, dim), comp)] - u[_rng(setup, lo, hi, z, comp)]) / _tab(
            setup, setup.DeltaU[dim], dim, lo[dim], hi[dim])

    if D == 2:
        omega[_rng(setup, lo, hi, z)] = d(1, 0) - d(0, 1)
    else:
        for i, j, k in ((0, 1, 2), (1, 2, 0), (2, 0, 1)):
            omega[_rng(setup, lo, hi, z, i)] = d(k, j) - d(j, k)
    return omega


def vorticity(u, setup):
    return vorticity_(scalarfield(setup) if setup.D == 2 else vectorfield(setup), u, setup)


def interpolate_u_p_(up, u, setup):
    """interpolate_u_p! (operators.jl:921-940)."""
    D, N = setup.D, setup.N
    T = setup.T
    for i in range(D):
        lo = tuple(1 if a == i else 0 for a in range(D))
        hi = tuple(N)
        z = _zero(D)
        up[_rng(setup, lo, hi, z, i)] = (u[_rng(setup, lo, hi, _sub(z, _e(D, i)), i)] + u[_rng(setup, lo, hi, z, i)]) / T(2)
    return up


def interpolate_u_p(u, setup):
    return interpolate_u_p_(vectorfield(setup), u, setup)


def interpolate_omega_p_(omegap, omega, setup):
    """interpolate_ω_p! (operators.jl:950-984): over Ip."""
    D = setup.D
    T = setup.T
    lo = tuple(s.start for s in setup.Ip)
    hi = tuple(s.stop for s in setup.Ip)
    z = _zero(D)
    if D == 2:
        omegap[_rng(setup, lo, hi, z)] = (omega[_rng(setup, lo, hi, (-1, -1))] + omega[_rng(setup, lo, hi, z)]) / T(2)
    else:
        for i in range(3):
            j, k = (i + 1) % 3, (i + 2) % 3
            off = tuple(-1 if a in (j, k) else 0 for a in range(3))
            omegap[_rng(setup, lo, hi, z, i)] = (omega[_rng(setup, lo, hi, off, i)] + omega[_rng(setup, lo, hi, z, i)]) / T(2)
    return omegap


def interpolate_omega_p(omega, setup):
    return interpolate_omega_p_(scalarfield(setup) if setup.D == 2 else vectorfield(setup), omega, setup)


def _delta_at(setup, u, i, j, lo, hi, off):
    """guarded δ(u, i, j, I + off) for I in lo:hi (operators.jl:179-188)."""
    D, T = setup.D, setup.T
    tol = 2 * np.finfo(T).eps
    ej = _e(D, j)
    if i == j:
        dij = _tab(setup, setup.Delta[j], j, lo[j] + off[j], hi[j] + off[j])
        right = u[_rng(setup, lo, hi, off, i)]
        left = u[_rng(setup, lo, hi, _sub(off, ej), i)]
    else:
        dij = _tab(setup, setup.DeltaU[j], j, lo[j] + off[j], hi[j] + off[j])
        right = u[_rng(setup, lo, hi, tuple(o + e for o, e in zip(off, ej)), i)]
        left = u[_rng(setup, lo, hi, off, i)]
    return np.where(dij > tol, (right - left) / dij, T(0))


def _delta_coll(setup, u, i, j, lo, hi):
    """δ_coll (operators.jl:883-893)."""
    D, T = setup.D, setup.T
    z = _zero(D)
    if i == j:
        return _delta_at(setup, u, i, j, lo, hi, z)
    ei, ej = _e(D, i), _e(D, j)
    return (
        _delta_at(setup, u, i, j, lo, hi, z)
        + _delta_at(setup, u, i, j, lo, hi, _sub(z, ei))
        + _delta_at(setup, u, i, j, lo, hi, _sub(z, ej))
        + _delta_at(setup, u, i, j, lo, hi, _sub(_sub(z, ei), ej))
    ) / T(4)


def qcrit_(q, u, setup):
    """qcrit! (operators.jl:989-1000): I = index + oneunit over Np, q = -tr(G G) / 2."""
    D, T = setup.D, setup.T
    lo = (1,) * D
    hi = tuple(1 + n for n in setup.Np)
    G = [[_delta_coll(setup, u, i, j, lo, hi) for j in range(D)] for i in range(D)]
    tr = None
    for i in range(D):
        d = None
        for k in range(D):
            t = G[i][k] * G[k][i]
            d = t if d is None else d + t
        tr = d if tr is None else tr + d
    q[_rng(setup, lo, hi, _zero(D))] = -tr / T(2)
    return q


def qcrit(u, setup):
    return qcrit_(scalarfield(setup), u, setup)


def kinetic_energy_(ke, u, setup, interpolate_first=False):
    """kinetic_energy! (operators.jl:1027-1055): over Ip."""
    D, T = setup.D, setup.T
    lo = tuple(s.start for s in setup.Ip)
    hi = tuple(s.stop for s in setup.Ip)
    z = _zero(D)
    k = np.zeros(tuple(h - l for l, h in zip(lo, hi)), dtype=T)
    for i in range(D):
        a = u[_rng(setup, lo, hi, z, i)]
        b = u[_rng(setup, lo, hi, _sub(z, _e(D, i)), i)]
        k = k + ((a + b) ** 2 if interpolate_first else a**2 + b**2)
    ke[_rng(setup, lo, hi, z)] = k / T(8 if interpolate_first else 4)
    return ke


def kinetic_energy(u, setup, interpolate_first=False):
    return kinetic_energy_(scalarfield(setup), u, setup, interpolate_first)


def total_kinetic_energy(u, setup, interpolate_first=False):
    """total_kinetic_energy (operators.jl:1061-1066)."""
    k = kinetic_energy(u, setup, interpolate_first)
    k = scalewithvolume_(k, setup)
    return k[setup.Ip].sum()


def energyspectrum(u, setup, kmax=None):
    """energyspectrum (src/spectral.jl:75-117, shells of spectral_stuff :36-66): ehat[kappa] = sum over
    kappa <= |k| < kappa + 1 of |u_hat|^2 / 2 with u_hat = rfft(u) / prod(Np), the modes whose
    conjugate is not stored counted twice.  Returns (kappa, ehat)."""
    import scipy.fft as sfft

    D, Np = setup.D, setup.Np
    kmax = min(Np) // 2 if kmax is None else kmax
    axes = tuple(range(D - 1, -1, -1))  # real transform along dim 0 (Julia dim 1)
    ntot = float(np.prod(Np))
    efield = 0
    for a in range(D):
        uhat = sfft.rfftn(np.ascontiguousarray(u[setup.Ip + (a,)]).astype(np.float64), axes=axes)
        efield = efield + np.abs(uhat) ** 2
    efield = efield / (2 * ntot**2)
    kx = np.arange(Np[0] // 2 + 1)
    k2 = (kx**2).reshape((-1,) + (1,) * (D - 1)).astype(np.int64)
    for a in range(1, D):
        ka = np.fft.fftfreq(Np[a], 1.0 / Np[a]).astype(np.int64)
        shape = [1] * D
        shape[a] = Np[a]
        k2 = k2 + (ka**2).reshape(shape)
    shell = np.floor(np.sqrt(k2)).astype(np.int64)
    shell = np.where((shell + 1) ** 2 <= k2, shell + 1, shell)
    shell = np.where(shell**2 > k2, shell - 1, shell)
    weight = np.where((kx > 0) & (kx < Np[0] // 2), 2.0, 1.0).reshape((-1,) + (1,) * (D - 1))
    w = efield * weight
    ehat = np.array([w[shell == kap].sum() for kap in range(kmax + 1)])
    return np.arange(kmax + 1), ehat


def spectral_energy_field(u, setup):
    """spectral_energy_field! (src/spectral.jl:75-87): sum over the components of |rfft(u_a)|^2 / (2 prod(Np)^2)
    on the RFFT-sized array (real transform along dim 1 of the reference)."""
    import scipy.fft as sfft

    D, Np, T = setup.D, setup.Np, setup.T
    axes = tuple(range(D - 1, -1, -1))
    efield = 0
    for a in range(D):
        uhat = sfft.rfftn(np.ascontiguousarray(u[setup.Ip + (a,)]), axes=axes)
        efield = efield + (uhat.real**2 + uhat.imag**2).astype(T)
    return efield / (2 * T(np.prod(Np)) ** 2)


def spectralsum(f):
    """spectralsum (src/spectral.jl:89-90): sum over an RFFT-sized array, the missing conjugate modes counted."""
    return f.sum() + f[1:-1].sum()


def squared_wavenumbers_dimensional(setup):
    """squared_wavenumbers_dimensional (src/spectral.jl:121-135): |2 pi k / L|^2 on the RFFT-sized array."""
    D, Np, T = setup.D, setup.Np, setup.T
    k2 = 0
    for a in range(D):
        L = T(setup.xlims[a][1] - setup.xlims[a][0])
        kint = np.arange(Np[0] // 2 + 1) if a == 0 else np.fft.fftfreq(Np[a], 1.0 / Np[a])
        shape = [1] * D
        shape[a] = -1
        k2 = k2 + ((T(2 * np.pi) * kint.astype(T) / L) ** 2).reshape(shape)
    return k2


def turbulence_statistics(u, setup, viscosity):
    """turbulence_statistics (src/spectral.jl:144-257): Kolmogorov scaling in 3-D, Kraichnan-Batchelor-Leith
    scaling in 2-D; uniform periodic grids only (assert_uniform_periodic, :10-23)."""
    from .grid import PeriodicBC

    T, D, Np = setup.T, setup.D, setup.Np
    for a in range(D):
        assert all(isinstance(b, PeriodicBC) for b in setup.bc_u[a]), "Turbulence statistics requires periodic boundary conditions"
        d = np.asarray(setup.Delta[a])
        assert np.all(np.abs(d - d[0]) <= np.sqrt(np.finfo(T).eps) * abs(d[0])), "Turbulence statistics requires a uniform grid"
    nu = T(viscosity)
    ef = spectral_energy_field(u, setup)
    k2 = squared_wavenumbers_dimensional(setup)
    e = spectralsum(ef)
    diss = 2 * nu * spectralsum(k2 * ef)
    kmax = T(np.pi) * min(Np[a] / (setup.xlims[a][1] - setup.xlims[a][0]) for a in range(D))
    if D == 3:
        uavg = np.sqrt(2 * e / 3)
        l_kol = (nu**3 / diss) ** T(1 / 4)
        l_tay = np.sqrt(15 * nu / diss) * uavg
        l_int = uavg**3 / diss
        return dict(e=e, uavg=uavg, diss=diss, l_int=l_int, l_tay=l_tay, l_kol=l_kol, t_int=l_int / uavg,
                    t_tay=l_tay / uavg, t_kol=np.sqrt(nu / diss), Re_int=l_int * uavg / nu, Re_tay=l_tay * uavg / nu,
                    kmax_eta=kmax * l_kol)
    uavg = np.sqrt(e)
    enstrophy = diss / (2 * nu)
    omega_rms = np.sqrt(2 * enstrophy)
    enstrophy_diss = 2 * nu * spectralsum(k2**2 * ef)
    l_kra = (nu**3 / enstrophy_diss) ** T(1 / 6)
    l_omega = uavg / omega_rms
    return dict(e=e, uavg=uavg, diss=diss, enstrophy=enstrophy, palinstrophy=enstrophy_diss / (2 * nu),
                enstrophy_diss=enstrophy_diss, omega_rms=omega_rms, l_kra=l_kra, l_omega=l_omega, t_eddy=1 / omega_rms,
                t_ens=enstrophy_diss ** (-T(1 / 3)), Re=uavg * l_omega / nu, kmax_etaK=kmax * l_kra)
