"""Oracle (test infrastructure): grid set-up.

Restates /root/reference/src/grid.jl:30-208 and the ghost padding / DOF offsets
of /root/reference/src/boundary_conditions.jl:1-107.

Index convention: the reference uses 1-based Julia indices; every array here is
0-based, i.e. Julia ``a[i]`` is ``a[i-1]``.  Fields are Fortran-ordered NumPy
arrays of shape ``N`` (scalar) or ``N + (D,)`` (vector) so that the memory
layout is exactly the reference's column-major layout
(src/initializers.jl:2,5).
"""

from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------
# Boundary-condition types (src/boundary_conditions.jl:1-39)
# ----------------------------------------------------------------------------
class PeriodicBC:
    code = 0

    def __repr__(self):
        return "PeriodicBC()"


class DirichletBC:
    """``u`` is None (no-slip), a tuple of D constants, a number (temperature)
    or a callable ``(dim, x..., t) -> value`` (src/boundary_conditions.jl:7-22)."""

    code = 1

    def __init__(self, u=None):
        self.u = u

    def __repr__(self):
        return f"DirichletBC({self.u!r})"


class SymmetricBC:
    code = 2

    def __repr__(self):
        return "SymmetricBC()"


class PressureBC:
    code = 3

    def __repr__(self):
        return "PressureBC()"


def offset_u(bc, isright, isnormal):
    """src/boundary_conditions.jl:83-92"""
    if isinstance(bc, PeriodicBC):
        return 1
    if isinstance(bc, (DirichletBC, SymmetricBC)):
        return 1 + int(isright) * int(isnormal)
    if isinstance(bc, PressureBC):
        return 1 + int(not isright) * int(not isnormal)
    raise TypeError(bc)


def offset_p(bc, isright):
    """src/boundary_conditions.jl:84-93"""
    if isinstance(bc, PressureBC):
        return 1 + int(not isright)
    return 1


def _padghost(bc, x, isright):
    """src/boundary_conditions.jl:45-64 (x is a 1-D array of dtype T)."""
    if isinstance(bc, PeriodicBC):
        if isright:
            dxa = x[1] - x[0]
            return np.append(x, x[-1] + dxa)
        dxb = x[-1] - x[-2]
        return np.insert(x, 0, x[0] - dxb)
    if isinstance(bc, DirichletBC):
        return np.append(x, x[-1]) if isright else np.insert(x, 0, x[0])
    if isinstance(bc, SymmetricBC):
        if isright:
            return np.append(x, x[-1] + (x[-1] - x[-2]))
        return np.insert(x, 0, x[0] - (x[1] - x[0]))
    if isinstance(bc, PressureBC):
        if isright:
            return np.append(x, x[-1])
        return np.insert(x, 0, [x[0], x[0]])
    raise TypeError(bc)


# ----------------------------------------------------------------------------
# 1-D grids (src/grid.jl:30-68)
# ----------------------------------------------------------------------------
def _sinpi(x):
    """sin(pi x) for float64 x, evaluated in extended precision then rounded
    (stand-in for Julia's correctly-rounded-ish ``sinpi``)."""
    x = np.asarray(x, dtype=np.float64)
    xl = x.astype(np.longdouble)
    return np.sin(_PI_LD * xl).astype(np.float64)


def _cospi(x):
    x = np.asarray(x, dtype=np.float64)
    xl = x.astype(np.longdouble)
    return np.cos(_PI_LD * xl).astype(np.float64)


# pi to long-double precision (x87 80-bit has 64-bit mantissa)
_PI_LD = np.longdouble("3.14159265358979323846264338327950288419716939937510")


def cosine_grid(a, b, N, dtype=np.float64):
    """src/grid.jl:30-34"""
    T = np.dtype(dtype).type
    i = np.arange(N + 1).astype(T)
    cp = _cospi((i / T(N)).astype(np.float64)).astype(T)
    return (T(a) + (T(b) - T(a)) * (1 - cp) / 2).astype(T)


def stretched_grid(a, b, N, s=1, dtype=np.float64):
    """src/grid.jl:51-58"""
    T = np.dtype(dtype).type
    if s <= 0:
        raise ValueError("The stretch factor must be positive")
    if np.isclose(s, 1):
        return np.linspace(a, b, N + 1, dtype=T)
    i = np.arange(N + 1)
    return (T(a) + (T(b) - T(a)) * (1 - T(s) ** i) / (1 - T(s) ** N)).astype(T)


def tanh_grid(a, b, N, gamma=1, dtype=np.float64):
    """src/grid.jl:64-68"""
    T = np.dtype(dtype).type
    x = np.linspace(0, 1, N + 1, dtype=T)
    g = T(gamma)
    return (T(a) + (T(b) - T(a)) * (1 + np.tanh(g * (2 * x - 1)) / np.tanh(g)) / 2).astype(T)


# ----------------------------------------------------------------------------
# Setup (src/grid.jl:91-208)
# ----------------------------------------------------------------------------
class Setup:
    """0-based restatement of the ``Setup`` NamedTuple.

    ``x``: tuple of D 1-D coordinate arrays (volume faces, without ghosts).
    ``boundary_conditions``: dict with key ``"u"`` (and optionally ``"temp"``):
    a D-tuple of ``(bc_left, bc_right)`` pairs.
    """

    def __init__(self, x, boundary_conditions):
        x = [np.array(xi) for xi in x]  # collect (grid.jl:93)
        T = x[0].dtype.type
        if T not in (np.float32, np.float64):
            x = [xi.astype(np.float64) for xi in x]
            T = np.float64
        x = [xi.astype(T) for xi in x]
        self.T = T
        self.dtype = np.dtype(T)
        self.D = D = len(x)
        self.xlims = tuple((xi.min(), xi.max()) for xi in x)  # grid.jl:94
        self.boundary_conditions = boundary_conditions
        bcu = boundary_conditions["u"]
        self.bc_u = bcu
        self.bc_temp = boundary_conditions.get("temp", None)

        # ghost padding, left first then right (grid.jl:105-109)
        for d in range(D):
            a, b = bcu[d]
            x[d] = _padghost(a, x[d], False)
            x[d] = _padghost(b, x[d], True)
        self.x = tuple(x)

        self.N = N = tuple(len(xi) - 1 for xi in x)  # grid.jl:112
        # DOF counts / ranges (grid.jl:115-145); slices are 0-based half-open
        self.Nu = tuple(
            tuple(
                N[b] - offset_u(bcu[b][0], False, a == b) - offset_u(bcu[b][1], True, a == b)
                for b in range(D)
            )
            for a in range(D)
        )
        self.Np = tuple(
            N[a] - offset_p(bcu[a][0], False) - offset_p(bcu[a][1], True) for a in range(D)
        )
        self.Iu = tuple(
            tuple(
                slice(
                    offset_u(bcu[b][0], False, a == b),
                    N[b] - offset_u(bcu[b][1], True, a == b),
                )
                for b in range(D)
            )
            for a in range(D)
        )
        self.Ip = tuple(
            slice(offset_p(bcu[a][0], False), N[a] - offset_p(bcu[a][1], True)) for a in range(D)
        )
        self.inside = tuple(slice(1, n - 1) for n in N)  # grid.jl:147-149

        # coordinates of velocity / pressure points (grid.jl:152-163)
        self.xp = tuple((xi[:-1] + xi[1:]) / 2 for xi in x)
        self.xu = tuple(
            tuple(x[b][1:] if a == b else (x[b][:-1] + x[b][1:]) / 2 for b in range(D))
            for a in range(D)
        )

        eps = np.finfo(T).eps
        # volume widths (grid.jl:167-171)
        self.Delta = tuple(np.maximum(np.diff(xi), eps).astype(T) for xi in x)
        # distance between pressure points (grid.jl:173-177)
        self.DeltaU = tuple(
            np.maximum(np.append(np.diff(self.xp[d]), self.Delta[d][-1] / 2), eps).astype(T)
            for d in range(D)
        )
        # interpolation weights (grid.jl:180-186)
        A1, A2 = [], []
        for j in range(D):
            Aj2 = (x[j][1 : N[j]] - self.xp[j][0 : N[j] - 1]) / self.DeltaU[j][0 : N[j] - 1]
            Aj1 = 1 - Aj2
            A1.append(np.insert(Aj1, 0, 1).astype(T))
            A2.append(np.append(Aj2, 1).astype(T))
        self.A1 = tuple(A1)
        self.A2 = tuple(A2)

    # convenience ------------------------------------------------------------
    def is_periodic(self, a):
        return isinstance(self.bc_u[a][0], PeriodicBC) and isinstance(self.bc_u[a][1], PeriodicBC)

    def is_wall(self, a):
        return isinstance(self.bc_u[a][0], DirichletBC) and isinstance(self.bc_u[a][1], DirichletBC)


def scalarfield(setup):
    """src/initializers.jl:2"""
    return np.zeros(setup.N, dtype=setup.dtype, order="F")


def vectorfield(setup):
    """src/initializers.jl:5"""
    return np.zeros(setup.N + (setup.D,), dtype=setup.dtype, order="F")
