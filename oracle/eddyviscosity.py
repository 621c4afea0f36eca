"""Oracle (test infrastructure): eddy-viscosity closures of /root/reference/src/eddyviscosity.jl
(3-D `eddy_viscosity_closure!` with the Smagorinsky / WALE / Vreman / QR models), restated with
NumPy slicing; tensors are dicts of scalar fields keyed like the reference's named tuples."""
from __future__ import annotations

import numpy as np

from . import scalarfield
from . import bc as _bc
from .grid import DirichletBC

NAMES = ("xx", "yx", "zx", "xy", "yy", "zy", "xz", "yz", "zz")  # tensorfield (initializers.jl:15-25)
AX = {"x": 0, "y": 1, "z": 2}


def tensorfield(setup):
    return {k: scalarfield(setup) for k in NAMES}


def _int(setup, off=(0, 0, 0)):
    return tuple(slice(1 + o, n - 1 + o) for n, o in zip(setup.N, off))


def _tab(setup, arr, dim, off=0):
    n = setup.N[dim]
    v = arr[1 + off : n - 1 + off]
    shape = [1, 1, 1]
    shape[dim] = len(v)
    return v.reshape(shape)


def _e(a):
    return tuple(1 if b == a else 0 for b in range(3))


def _neg(*axes):
    return tuple(-1 if b in axes else 0 for b in range(3))


def gradient_tensor_(G, u, setup):
    """gradient_tensor_kernel! (eddyviscosity.jl:57-74): G.ab = ∂u_a/∂x_b on the staggered positions."""
    for a in range(3):
        for b in range(3):
            name = "xyz"[a] + "xyz"[b]
            if a == b:
                G[name][_int(setup)] = (u[_int(setup) + (a,)] - u[_int(setup, _neg(a)) + (a,)]) / _tab(setup, setup.Delta[a], a)
            else:
                G[name][_int(setup)] = (u[_int(setup, _e(b)) + (a,)] - u[_int(setup) + (a,)]) / _tab(setup, setup.DeltaU[b], b)
    return G


def zero_out_wall_(p, setup):
    """zero_out_wall! (eddyviscosity.jl:344-353)."""
    for i in range(setup.D):
        bc = setup.bc_u[i]
        idx = [slice(None)] * setup.D
        if isinstance(bc[0], DirichletBC):
            idx[i] = 0
            p[tuple(idx)] = 0
        if isinstance(bc[1], DirichletBC):
            idx[i] = p.shape[i] - 1
            p[tuple(idx)] = 0


def collocate_tensor(G, setup):
    """collocate_tensor (eddyviscosity.jl:235-250): M[a][b] at the cell centres of the interior."""
    T = setup.T
    M = [[None] * 3 for _ in range(3)]
    for a in range(3):
        for b in range(3):
            s = G["xyz"[a] + "xyz"[b]]
            if a == b:
                M[a][b] = s[_int(setup)]
            else:
                lo, hi = min(a, b), max(a, b)  # offsets -e_lo, -e_hi, -e_lo-e_hi in the reference's order
                M[a][b] = (s[_int(setup)] + s[_int(setup, _neg(lo))] + s[_int(setup, _neg(hi))] + s[_int(setup, _neg(a, b))]) / T(4)
    return M


def _dot(A, B):
    s = None
    for b in range(3):  # column-major accumulation like sum over an SMatrix
        for a in range(3):
            t = A[a][b] * B[a][b]
            s = t if s is None else s + t
    return s


def _matmul(A, B):
    return [[A[a][0] * B[0][b] + A[a][1] * B[1][b] + A[a][2] * B[2][b] for b in range(3)] for a in range(3)]


def _tr(A):
    return A[0][0] + A[1][1] + A[2][2]


def gridsize_vol(setup):
    return np.cbrt(_tab(setup, setup.Delta[0], 0) * _tab(setup, setup.Delta[1], 1) * _tab(setup, setup.Delta[2], 2))


def eddy_viscosity_(model, C, visc, G, setup):
    """eddy_viscosity_kernel! (eddyviscosity.jl:265-340)."""
    T = setup.T
    C = T(C)
    M = collocate_tensor(G, setup)
    d = gridsize_vol(setup).astype(T)
    S = [[(M[a][b] + M[b][a]) / T(2) for b in range(3)] for a in range(3)]
    if model == "smagorinsky":
        val = (C * d) ** 2 * np.sqrt(T(2) * _dot(S, S))
    elif model == "wale":
        G2 = _matmul(M, M)
        tr = _tr(G2)
        Sd = [[(G2[a][b] + G2[b][a]) / T(2) - (tr / T(3) if a == b else T(0)) for b in range(3)] for a in range(3)]
        sdsd, ss = _dot(Sd, Sd), _dot(S, S)
        val = (C * d) ** 2 * sdsd ** T(3 / 2) / (ss ** T(5 / 2) + sdsd ** T(5 / 4) + np.finfo(T).eps)
    elif model == "qr":
        SS = _matmul(S, S)
        QS = _tr(SS) / T(2)
        RS = _tr(_matmul(SS, S)) / T(3)
        val = (C * d) ** 2 * np.abs(RS) / (QS + np.finfo(T).eps)
    elif model == "vreman":
        d1 = _tab(setup, setup.Delta[0], 0) ** 2
        d2 = _tab(setup, setup.Delta[1], 1) ** 2
        d3 = _tab(setup, setup.Delta[2], 2) ** 2
        a11, a21, a31 = M[0][0], M[0][1], M[0][2]
        a12, a22, a32 = M[1][0], M[1][1], M[1][2]
        a13, a23, a33 = M[2][0], M[2][1], M[2][2]
        b11 = d1 * a11 * a11 + d2 * a21 * a21 + d3 * a31 * a31
        b12 = d1 * a11 * a12 + d2 * a21 * a22 + d3 * a31 * a32
        b13 = d1 * a11 * a13 + d2 * a21 * a23 + d3 * a31 * a33
        b22 = d1 * a12 * a12 + d2 * a22 * a22 + d3 * a32 * a32
        b23 = d1 * a12 * a13 + d2 * a22 * a23 + d3 * a32 * a33
        b33 = d1 * a13 * a13 + d2 * a23 * a23 + d3 * a33 * a33
        abeta = a11**2 + a12**2 + a13**2 + a21**2 + a22**2 + a23**2 + a31**2 + a32**2 + a33**2
        bbeta = b11 * b22 - b12**2 + b11 * b33 - b13**2 + b22 * b33 - b23**2
        with np.errstate(divide="ignore", invalid="ignore"):
            value = C**2 * np.sqrt(bbeta / abeta)
        tol = 1e-12 if T is np.float64 else np.float32(1e-6)
        val = np.where(abeta > tol, value, T(0))
    else:
        raise ValueError(model)
    visc[_int(setup)] = val.astype(T)
    return visc


def symmetrize_(G):
    """symmetrize! (eddyviscosity.jl:81-86): whole arrays."""
    T = G["xy"].dtype.type
    G["xy"][...] = (G["xy"] + G["yx"]) / T(2)
    G["xz"][...] = (G["xz"] + G["zx"]) / T(2)
    G["yz"][...] = (G["yz"] + G["zy"]) / T(2)


def apply_eddy_viscosity_(sig, visc, setup):
    """apply_eddy_viscosity_kernel! 3-D (eddyviscosity.jl:144-181)."""
    T = setup.T
    w = []
    for dim in range(3):
        da, db = _tab(setup, setup.Delta[dim], dim), _tab(setup, setup.Delta[dim], dim, 1)
        w.append((db / (da + db), da / (da + db)))
    (ax, bx), (ay, by), (az, bz) = w
    v = lambda *axes: visc[_int(setup, tuple(1 if b in axes else 0 for b in range(3)))]
    visc_xy = ax * ay * v() + bx * ay * v(0) + ax * by * v(1) + bx * by * v(0, 1)
    visc_xz = ax * az * v() + bx * az * v(0) + ax * bz * v(2) + bx * bz * v(0, 2)
    visc_yz = ay * az * v() + by * az * v(1) + ay * bz * v(2) + by * bz * v(1, 2)
    I = _int(setup)
    for k in ("xx", "yy", "zz"):
        sig[k][I] = T(-2) * v() * sig[k][I]
    sig["xy"][I] = T(-2) * visc_xy * sig["xy"][I]
    sig["xz"][I] = T(-2) * visc_xz * sig["xz"][I]
    sig["yz"][I] = T(-2) * visc_yz * sig["yz"][I]


def divoftensor_(f, sig, setup):
    """divoftensor_kernel! 3-D (eddyviscosity.jl:200-219): f[I, a] -= sum_b d(sigma_ab)/dx_b."""
    I = _int(setup)
    dp = [_tab(setup, setup.Delta[a], a) for a in range(3)]
    du = [_tab(setup, setup.DeltaU[a], a) for a in range(3)]
    s = lambda name, off=(0, 0, 0): sig[name][_int(setup, off)]
    f[I + (0,)] -= (s("xx", _e(0)) - s("xx")) / du[0] + (s("xy") - s("xy", _neg(1))) / dp[1] + (s("xz") - s("xz", _neg(2))) / dp[2]
    f[I + (1,)] -= (s("xy") - s("xy", _neg(0))) / dp[0] + (s("yy", _e(1)) - s("yy")) / du[1] + (s("yz") - s("yz", _neg(2))) / dp[2]
    f[I + (2,)] -= (s("xz") - s("xz", _neg(0))) / dp[0] + (s("yz") - s("yz", _neg(1))) / dp[1] + (s("zz", _e(2)) - s("zz")) / du[2]
    return f


def eddy_viscosity_closure_(model, C, f, u, setup, cache=None):
    """eddy_viscosity_closure! (eddyviscosity.jl:356-378); model in {smagorinsky, wale, vreman, qr}."""
    T = setup.T
    cache = cache or dict(visc=scalarfield(setup), G=tensorfield(setup))
    visc, G = cache["visc"], cache["G"]
    visc[...] = 0
    for g in G.values():
        g[...] = 0
    gradient_tensor_(G, u, setup)
    for g in G.values():
        zero_out_wall_(g, setup)
        _bc.apply_bc_p(g, T(0), setup)
    eddy_viscosity_(model, C, visc, G, setup)
    zero_out_wall_(visc, setup)
    _bc.apply_bc_p(visc, T(0), setup)
    symmetrize_(G)
    apply_eddy_viscosity_(G, visc, setup)
    for g in G.values():
        zero_out_wall_(g, setup)
        _bc.apply_bc_p(g, T(0), setup)
    divoftensor_(f, G, setup)
    return f
