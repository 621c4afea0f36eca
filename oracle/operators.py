"""Oracle (test infrastructure): matrix-free staggered-grid operators.

Restates /root/reference/src/operators.jl (forward operators only) with NumPy
slicing.  Every kernel of the reference runs over ``I in 2:N-1`` in each
dimension (``apply!``, operators.jl:68-78); here that is the 0-based interior
``1:N-1`` and a point function evaluated at ``I + off`` becomes a shifted view.
Evaluation order (parentheses) follows the Julia source so that the IEEE
results are those of the reference's CPU path (no FMA contraction there).
"""

from __future__ import annotations

import numpy as np

# "numpy": the slicing restatement below (default, what the parity tests compare against);
# "c": the stencil operators run through oracle/c (cell-by-cell C/OpenMP restatement, the
# multithreaded CPU baseline of bench.py).  Both give the same IEEE results.
BACKEND = "numpy"


def _c():
    from . import ckernels

    return ckernels


# ---------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------
def _e(D, j):
    return tuple(1 if a == j else 0 for a in range(D))


def _add(a, b):
    return tuple(x + y for x, y in zip(a, b))


def _sub(a, b):
    return tuple(x - y for x, y in zip(a, b))


def _zero(D):
    return (0,) * D


def _sl(setup, off, comp=None):
    s = tuple(slice(1 + o, n - 1 + o) for n, o in zip(setup.N, off))
    return s if comp is None else s + (comp,)


def _m(setup, arr, dim, off=0):
    """1-D metric table over the interior (shifted by ``off``), broadcast along ``dim``."""
    n = setup.N[dim]
    v = arr[1 + off : n - 1 + off]
    shape = [1] * setup.D
    shape[dim] = len(v)
    return v.reshape(shape)


# ---------------------------------------------------------------------------
# point functions
# ---------------------------------------------------------------------------
def delta_u(setup, u, i, j, off):
    """Guarded velocity derivative ``δ(setup,u,i,j,I+off)`` (operators.jl:179-188)."""
    D = setup.D
    T = setup.T
    tol = 2 * np.finfo(T).eps
    ej = _e(D, j)
    if i == j:
        dij = _m(setup, setup.Delta[j], j, off[j])
        right = u[_sl(setup, off, i)]
        left = u[_sl(setup, _sub(off, ej), i)]
    else:
        dij = _m(setup, setup.DeltaU[j], j, off[j])
        right = u[_sl(setup, _add(off, ej), i)]
        left = u[_sl(setup, off, i)]
    du = (right - left) / dij
    return np.where(dij > tol, du, T(0))


def delta_p(setup, p, i, off):
    """Scalar derivative landing on the face (operators.jl:192)."""
    ei = _e(setup.D, i)
    return (p[_sl(setup, _add(off, ei))] - p[_sl(setup, off)]) / _m(
        setup, setup.DeltaU[i], i, off[i]
    )


def avg(setup, phi, off, i):
    """Linear interpolation of a scalar to the face in direction i (operators.jl:197-200)."""
    ei = _e(setup.D, i)
    d0 = _m(setup, setup.Delta[i], i, off[i])
    d1 = _m(setup, setup.Delta[i], i, off[i] + 1)
    return (d1 * phi[_sl(setup, off)] + d0 * phi[_sl(setup, _add(off, ei))]) / (d0 + d1)


def convstress(setup, u, i, j, off):
    """operators.jl:378-393"""
    D = setup.D
    T = setup.T
    ei, ej = _e(D, i), _e(D, j)
    if i == j:
        uij = (u[_sl(setup, _sub(off, ej), i)] + u[_sl(setup, off, i)]) / T(2)
        uji = uij
    else:
        uij = (u[_sl(setup, off, i)] + u[_sl(setup, _add(off, ej), i)]) / T(2)
        uji = _m(setup, setup.A2[i], i, off[i]) * u[_sl(setup, off, j)] + _m(
            setup, setup.A1[i], i, off[i] + 1
        ) * u[_sl(setup, _add(off, ei), j)]
    return uij * uji


def diffstress(setup, u, visc, i, j, off):
    """operators.jl:556"""
    return (-setup.T(visc)) * delta_u(setup, u, i, j, off)


def convdiffstress(setup, u, visc, i, j, off):
    """operators.jl:627-628"""
    return convstress(setup, u, i, j, off) + diffstress(setup, u, visc, i, j, off)


def tensordivergence(setup, sigma, i, j):
    """operators.jl:395-401, evaluated over the interior. ``sigma(i, j, off)``."""
    D = setup.D
    ej = _e(D, j)
    z = _zero(D)
    if i == j:
        dd = _m(setup, setup.DeltaU[j], j)
        sa = sigma(i, j, z)  # left(I, j, false) = I
        sb = sigma(i, j, ej)  # right(I, j, true)
    else:
        dd = _m(setup, setup.Delta[j], j)
        sa = sigma(i, j, _sub(z, ej))  # left(I, j, true)
        sb = sigma(i, j, z)
    return -(sb - sa) / dd


def _contract_tensor_add(setup, f, term):
    """contract_tensor_add! (operators.jl:129-142): f[I,i] += (t_i1 + t_i2) (+ t_i3)."""
    D = setup.D
    z = _zero(D)
    for i in range(D):
        s = term(i, 0) + term(i, 1)
        if D == 3:
            s = s + term(i, 2)
        f[_sl(setup, z, i)] += s
    return f


# ---------------------------------------------------------------------------
# operators
# ---------------------------------------------------------------------------
def convection_(f, u, setup):
    """convection! (operators.jl:407-416): adds to f."""
    if BACKEND == "c":
        return _c().convdiff_add(f, u, setup, 0.0, conv=True, diff=False)
    sig = lambda i, j, off: convstress(setup, u, i, j, off)
    return _contract_tensor_add(setup, f, lambda i, j: tensordivergence(setup, sig, i, j))


def diffusion_(f, u, setup, visc):
    """diffusion! (operators.jl:545-554): adds to f."""
    if BACKEND == "c":
        return _c().convdiff_add(f, u, setup, visc, conv=False, diff=True)
    sig = lambda i, j, off: diffstress(setup, u, visc, i, j, off)
    return _contract_tensor_add(setup, f, lambda i, j: tensordivergence(setup, sig, i, j))


def convectiondiffusion_(f, u, setup, visc):
    """convectiondiffusion! (operators.jl:616-628): adds to f."""
    if BACKEND == "c":
        return _c().convdiff_add(f, u, setup, visc)
    sig = lambda i, j, off: convdiffstress(setup, u, visc, i, j, off)
    return _contract_tensor_add(setup, f, lambda i, j: tensordivergence(setup, sig, i, j))


def divergence_(div, u, setup):
    """divergence! (operators.jl:242-247, contract_vector! :90-100)."""
    if BACKEND == "c":
        return _c().divergence(div, u, setup)
    D = setup.D
    z = _zero(D)
    s = delta_u(setup, u, 0, 0, z) + delta_u(setup, u, 1, 1, z)
    if D == 3:
        s = s + delta_u(setup, u, 2, 2, z)
    div[_sl(setup, z)] = s
    return div


def scalewithvolume_(p, setup):
    """scalewithvolume! (operators.jl:218-231): WHOLE array, p *= (Δx*Δy)*Δz."""
    if BACKEND == "c":
        return _c().scalewithvolume(p, setup)
    D = setup.D
    dx = setup.Delta[0].reshape((-1,) + (1,) * (D - 1))
    dy = setup.Delta[1].reshape((1, -1) + (1,) * (D - 2))
    w = dx * dy
    if D == 3:
        w = w * setup.Delta[2].reshape(1, 1, -1)
    p *= w
    return p


def pressuregradient_(G, p, setup):
    """pressuregradient! (operators.jl:277-280, expand_scalar!)."""
    z = _zero(setup.D)
    for i in range(setup.D):
        G[_sl(setup, z, i)] = delta_p(setup, p, i, z)
    return G


def applypressure_(u, p, setup):
    """applypressure! (operators.jl:306-311): u[I,i] += -δ(p,i,I)."""
    if BACKEND == "c":
        return _c().applypressure(u, p, setup)
    z = _zero(setup.D)
    for i in range(setup.D):
        u[_sl(setup, z, i)] += -delta_p(setup, p, i, z)
    return u


def laplacian_(L, p, setup):
    """laplacian! (operators.jl:320-359); only used to manufacture test problems."""
    from .grid import DirichletBC, PressureBC

    D = setup.D
    T = setup.T
    L[...] = 0
    Ip = setup.Ip
    Np = setup.Np
    # restrict to the pressure DOFs (ndrange = Np, offset I0)
    def view(a, off):
        return a[tuple(slice(s.start + o, s.stop + o) for s, o in zip(Ip, off))]

    def met(arr, dim, off=0):
        v = arr[Ip[dim].start + off : Ip[dim].stop + off]
        shape = [1] * D
        shape[dim] = len(v)
        return v.reshape(shape)

    omega = met(setup.Delta[0], 0)
    for a in range(1, D):
        omega = omega * met(setup.Delta[a], a)
    z = _zero(D)
    for a in range(D):
        ea = _e(D, a)
        bc = setup.bc_u[a]
        pc = view(p, z)
        pr = view(p, ea)
        pl = view(p, _sub(z, ea))
        du = met(setup.DeltaU[a], a)
        dum = met(setup.DeltaU[a], a, -1)
        right = (pr - pc) / du
        left = (pc - pl) / dum
        # index masks along dim a (first / last pressure DOF)
        idx = np.arange(Np[a])
        shape = [1] * D
        shape[a] = Np[a]
        first = (idx == 0).reshape(shape)
        last = (idx == Np[a] - 1).reshape(shape)
        general = right - left
        val = general
        if isinstance(bc[1], DirichletBC):
            val = np.where(last, -left, val)
        if isinstance(bc[0], DirichletBC):
            val = np.where(first, right, val)
        if isinstance(bc[1], PressureBC):
            val = np.where(last, (-pc) / du - left, val)
        if isinstance(bc[0], PressureBC):
            val = np.where(first, right - pc / dum, val)
        Lv = view(L, z)
        Lv += omega / met(setup.Delta[a], a) * val
    return L


# ---------------------------------------------------------------------------
# Boussinesq terms
# ---------------------------------------------------------------------------
def applygravity_(f, temp, setup, gdir, gravity):
    """applygravity! (operators.jl:802-829)."""
    z = _zero(setup.D)
    f[_sl(setup, z, gdir)] += setup.T(gravity) * avg(setup, temp, z, gdir)
    return f


def convection_diffusion_temp_(c, u, temp, setup, conductivity):
    """convection_diffusion_temp! (operators.jl:669-681), contract_vector_add!."""
    D = setup.D
    T = setup.T
    z = _zero(D)
    cond = T(conductivity)

    def term(i):
        ei = _e(D, i)
        mi = _sub(z, ei)
        dT1 = (temp[_sl(setup, z)] - temp[_sl(setup, mi)]) / _m(setup, setup.DeltaU[i], i, -1)
        dT2 = (temp[_sl(setup, ei)] - temp[_sl(setup, z)]) / _m(setup, setup.DeltaU[i], i)
        uT1 = u[_sl(setup, mi, i)] * avg(setup, temp, mi, i)
        uT2 = u[_sl(setup, z, i)] * avg(setup, temp, z, i)
        return (-(uT2 - uT1) + cond * (dT2 - dT1)) / _m(setup, setup.Delta[i], i)

    s = term(0) + term(1)
    if D == 3:
        s = s + term(2)
    c[_sl(setup, z)] += s
    return c


def _delta_coll(setup, u, i, j):
    """δ_coll (operators.jl:883-893) over the interior."""
    D = setup.D
    z = _zero(D)
    if i == j:
        return delta_u(setup, u, i, j, z)
    ei, ej = _e(D, i), _e(D, j)
    return (
        delta_u(setup, u, i, j, z)
        + delta_u(setup, u, i, j, _sub(z, ei))
        + delta_u(setup, u, i, j, _sub(z, ej))
        + delta_u(setup, u, i, j, _sub(_sub(z, ei), ej))
    ) / setup.T(4)


def dissipation_(diss, u, setup, coeff):
    """dissipation! (operators.jl:772-781) with ∇_coll (:894-910); dot(G,G) is
    accumulated in column-major element order."""
    D = setup.D
    z = _zero(D)
    acc = None
    for j in range(D):  # column
        for i in range(D):  # row
            g = _delta_coll(setup, u, i, j)
            acc = g * g if acc is None else acc + g * g
    diss[_sl(setup, z)] += setup.T(coeff) * acc
    return diss


# ---------------------------------------------------------------------------
# right-hand sides (src/solver.jl:2-35)
# ---------------------------------------------------------------------------
def navierstokes_(force, state, t, setup, viscosity):
    """navierstokes! (solver.jl:2-6)."""
    force["u"][...] = 0
    convectiondiffusion_(force["u"], state["u"], setup, viscosity)
    return force


def boussinesq_(force, state, t, setup, viscosity, conductivity, gdir, gravity, dodissipation):
    """boussinesq! (solver.jl:16-35)."""
    u, temp = state["u"], state["temp"]
    force["u"][...] = 0
    force["temp"][...] = 0
    convectiondiffusion_(force["u"], u, setup, viscosity)
    applygravity_(force["u"], temp, setup, gdir, gravity)
    convection_diffusion_temp_(force["temp"], u, temp, setup, conductivity)
    if dodissipation:
        dissipation_(force["temp"], u, setup, viscosity)
    return force


# ---------------------------------------------------------------------------
# CFL (src/solver.jl:157-214)
# ---------------------------------------------------------------------------
def propose_timestep(state, setup, viscosity, conductivity=None):
    """propose_timestep(navierstokes!/boussinesq!, ...) (solver.jl:157-214)."""
    D = setup.D
    T = setup.T
    u = state["u"]
    # convective (:171-186)
    dt_conv = None
    for a in range(D):
        Iu = setup.Iu[a]
        da = setup.DeltaU[a][Iu[a]]
        shape = [1] * D
        shape[a] = len(da)
        ua = u[tuple(Iu) + (a,)]
        with np.errstate(divide="ignore"):
            v = np.min(da.reshape(shape) / np.abs(ua))
        dt_conv = v if dt_conv is None else min(dt_conv, v)
    # diffusive (:157-166)
    dt_diff = min(
        T(np.min(setup.DeltaU[a][setup.Iu[a][a]])) ** 2 / T(viscosity) / (2 * D) for a in range(D)
    )
    dt = min(dt_conv, dt_diff)
    if conductivity is not None:  # :188-197
        dt_temp = min(
            T(np.min(setup.Delta[a][setup.Ip[a]])) ** 2 / T(conductivity) / (2 * D)
            for a in range(D)
        )
        dt = min(dt, dt_temp)
    return T(dt)
