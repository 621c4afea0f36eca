"""Oracle (test infrastructure): ghost-cell boundary fills.

Restates /root/reference/src/boundary_conditions.jl:101-107 (``boundary``),
:163-171 / :200-208 / :237-245 (drivers) and the per-type methods
:273-285, :303-315, :335-336 (periodic), :341-372, :385, :388-402 (Dirichlet),
:411-425, :442-450, :463-464 (symmetric), :469-479, :494-499, :509-510
(pressure).  0-based indices; user BC functions receive a 0-based ``dim``.
"""

from __future__ import annotations

import numbers

import numpy as np

from .grid import DirichletBC, PeriodicBC, PressureBC, SymmetricBC


def _plane(setup, beta, I, isright):
    """``boundary(β, N, I, isright)`` (boundary_conditions.jl:101-107): index of
    the plane just outside the DOF range ``I`` (tuple of slices)."""
    return I[beta].stop if isright else I[beta].start - 1


def _idx(D, beta, i, rest=()):
    return tuple(i if a == beta else slice(None) for a in range(D)) + tuple(rest)


def _coords(D, xs, beta, i):
    """Broadcastable coordinate arrays over the plane ``beta == i``."""
    out = []
    for g in range(D):
        c = xs[g][i : i + 1] if g == beta else xs[g]
        shape = [1] * D
        shape[g] = len(c)
        out.append(c.reshape(shape))
    return out


# ----------------------------------------------------------------------------
def apply_bc_u(u, t, setup, dudt=False):
    """In-place ``apply_bc_u!`` (boundary_conditions.jl:163-171): dims in order,
    left side then right side."""
    D = setup.D
    for beta in range(D):
        for isright in (False, True):
            _bc_u_side(setup.bc_u[beta][int(isright)], u, beta, t, setup, isright, dudt)
    return u


def _bc_u_side(bc, u, beta, t, setup, isright, dudt):
    D, N = setup.D, setup.N
    T = setup.T
    if isinstance(bc, PeriodicBC):  # :273-285
        if isright:
            return
        ia = _plane(setup, beta, setup.Ip, False)
        ib = _plane(setup, beta, setup.Ip, True)
        u[_idx(D, beta, ia)] = u[_idx(D, beta, ib - 1)].copy()
        u[_idx(D, beta, ib)] = u[_idx(D, beta, ia + 1)].copy()
    elif isinstance(bc, DirichletBC):  # :341-372
        for a in range(D):
            i = _plane(setup, beta, setup.Iu[a], isright)
            if bc.u is None:
                val = T(0)
            elif isinstance(bc.u, tuple):
                val = T(0) if dudt else T(bc.u[a])
            else:
                xs = _coords(D, setup.xu[a], beta, i)
                if dudt:
                    h = T(np.sqrt(np.finfo(T).eps) / 2)
                    val = (bc.u(a, *xs, T(t) + h) - bc.u(a, *xs, T(t) - h)) / (2 * h)
                else:
                    val = bc.u(a, *xs, T(t))
                val = np.asarray(val, dtype=T)
            tgt = u[_idx(D, beta, slice(i, i + 1), (a,))]
            tgt[...] = val
    elif isinstance(bc, SymmetricBC):  # :411-425
        for a in range(D):
            i = _plane(setup, beta, setup.Iu[a], isright)
            if a == beta:
                u[_idx(D, beta, i, (a,))] = 0
            else:
                j = i - 1 if isright else i + 1
                u[_idx(D, beta, i, (a,))] = u[_idx(D, beta, j, (a,))].copy()
    elif isinstance(bc, PressureBC):  # :469-479
        for a in range(D):
            i = _plane(setup, beta, setup.Iu[a], isright)
            j = i - 1 if isright else i + 1
            u[_idx(D, beta, i, (a,))] = u[_idx(D, beta, j, (a,))].copy()
    else:
        raise TypeError(bc)


# ----------------------------------------------------------------------------
def apply_bc_p(p, t, setup):
    """In-place ``apply_bc_p!`` (boundary_conditions.jl:200-208)."""
    for beta in range(setup.D):
        for isright in (False, True):
            _bc_p_side(setup.bc_u[beta][int(isright)], p, beta, setup, isright)
    return p


def _bc_p_side(bc, p, beta, setup, isright):
    D = setup.D
    if isinstance(bc, PeriodicBC):  # :303-315
        if isright:
            return
        ia = _plane(setup, beta, setup.Ip, False)
        ib = _plane(setup, beta, setup.Ip, True)
        p[_idx(D, beta, ia)] = p[_idx(D, beta, ib - 1)].copy()
        p[_idx(D, beta, ib)] = p[_idx(D, beta, ia + 1)].copy()
    elif isinstance(bc, DirichletBC):  # :385 (no-op)
        return
    elif isinstance(bc, SymmetricBC):  # :442-450
        i = _plane(setup, beta, setup.Ip, isright)
        j = i - 1 if isright else i + 1
        p[_idx(D, beta, i)] = p[_idx(D, beta, j)].copy()
    elif isinstance(bc, PressureBC):  # :494-499
        i = _plane(setup, beta, setup.Ip, isright)
        p[_idx(D, beta, i)] = 0
    else:
        raise TypeError(bc)


# ----------------------------------------------------------------------------
def apply_bc_temp(temp, t, setup):
    """In-place ``apply_bc_temp!`` (boundary_conditions.jl:237-245)."""
    for beta in range(setup.D):
        for isright in (False, True):
            bc = setup.bc_temp[beta][int(isright)]
            if isinstance(bc, DirichletBC):  # :388-402
                D, T = setup.D, setup.T
                i = _plane(setup, beta, setup.Ip, isright)
                if bc.u is None:
                    val = T(0)
                elif isinstance(bc.u, numbers.Number):
                    val = T(bc.u)
                else:
                    xs = _coords(D, setup.xp, beta, i)
                    val = np.asarray(bc.u(*xs, T(t)), dtype=T)
                tgt = temp[_idx(D, beta, slice(i, i + 1))]
                tgt[...] = val
            elif isinstance(bc, PressureBC):  # :509-510 -> symmetric
                _bc_p_side(SymmetricBC(), temp, beta, setup, isright)
            else:  # periodic :335-336, symmetric :463-464
                _bc_p_side(bc, temp, beta, setup, isright)
    return temp
