"""Host-side problem definition: the Python mirror of the reference's ``Setup``
(src/grid.jl:91-208), BC types (src/boundary_conditions.jl:1-93) and 1-D grid helpers
(src/grid.jl:30-68).  Only O(N) host arithmetic lives here; the tables are handed to
libinsb200 as data (include/insb200.h: insb200_setup_t).

Indices are 0-based: Julia ``a[i]`` is ``a[i-1]``; ranges are half-open ``(lo, hi)``.
User callbacks (``ufunc(dim, x...)``, ``DirichletBC`` functions) get a 0-based ``dim``.
"""

from __future__ import annotations

import numbers

import numpy as np

from . import _lib


class PeriodicBC:
    code = _lib.BC_PERIODIC

    def __repr__(self):
        return "PeriodicBC()"


class DirichletBC:
    """``u``: None (no-slip), a tuple of D constants, a number (temperature BC), or a function
    ``u(dim, x..., t)`` (src/boundary_conditions.jl:7-20).  ``time_dependent`` declares whether a function
    depends on ``t``: True -> the steppers re-evaluate it at every stage time, False -> evaluated once;
    None (default) -> treated as time dependent as soon as two evaluations at different times differ
    (checked when the plan is built and again at the start of every step)."""

    code = _lib.BC_DIRICHLET

    def __init__(self, u=None, time_dependent=None):
        self.u = u
        self.time_dependent = time_dependent

    def __repr__(self):
        return f"DirichletBC({self.u!r})"


class SymmetricBC:
    code = _lib.BC_SYMMETRIC

    def __repr__(self):
        return "SymmetricBC()"


class PressureBC:
    code = _lib.BC_PRESSURE

    def __repr__(self):
        return "PressureBC()"


def _noff_u(code, right, normal):
    if code == _lib.BC_PERIODIC:
        return 1
    if code in (_lib.BC_DIRICHLET, _lib.BC_SYMMETRIC):
        return 2 if (right and normal) else 1
    return 2 if (not right and not normal) else 1


def _noff_p(code, right):
    return 2 if (code == _lib.BC_PRESSURE and not right) else 1


def _sincospi_ld(x):
    xl = np.asarray(x, dtype=np.float64).astype(np.longdouble)
    pi = np.longdouble("3.14159265358979323846264338327950288419716939937510")
    return np.sin(pi * xl), np.cos(pi * xl)


def cosine_grid(a, b, N, dtype=np.float64):
    T = np.dtype(dtype).type
    i = np.arange(N + 1).astype(T)
    _, c = _sincospi_ld((i / T(N)).astype(np.float64))
    return (T(a) + (T(b) - T(a)) * (1 - c.astype(T)) / 2).astype(T)


def stretched_grid(a, b, N, s=1, dtype=np.float64):
    T = np.dtype(dtype).type
    if s <= 0:
        raise ValueError("The stretch factor must be positive")
    if np.isclose(s, 1):
        return np.linspace(a, b, N + 1, dtype=T)
    i = np.arange(N + 1)
    return (T(a) + (T(b) - T(a)) * (1 - T(s) ** i) / (1 - T(s) ** N)).astype(T)


def tanh_grid(a, b, N, gamma=1, dtype=np.float64):
    T = np.dtype(dtype).type
    s = np.linspace(0, 1, N + 1, dtype=T)
    g = T(gamma)
    return (T(a) + (T(b) - T(a)) * (1 + np.tanh(g * (2 * s - 1)) / np.tanh(g)) / 2).astype(T)


class Setup:
    """``Setup(x=(x1, x2[, x3]), boundary_conditions={"u": ((l, r), ...)[, "temp": ...]})``.

    Holds the same fields as the reference NamedTuple (N, Nu, Np, Iu, Ip, x, xu, xp, Δ, Δu,
    A_stag, xlims, boundary_conditions); plans (device tables + workspaces) are created on
    demand by :mod:`insb200.api` and cached per psolver kind.
    """

    def __init__(self, x, boundary_conditions, device=None, rank=0, world=1, nccl_id=None):
        """``rank``/``world``/``nccl_id`` describe the z-slab decomposition (one process per
        GPU): ``x`` and the BCs are always those of the GLOBAL grid; this rank then owns the
        z planes ``[rank*nzl, rank*nzl + nzl + 2)`` of every ghosted array (``N_local``)."""
        xs = [np.array(v) for v in x]
        T = xs[0].dtype.type if xs[0].dtype.type in (np.float32, np.float64) else np.float64
        xs = [v.astype(T) for v in xs]
        self.T = T
        self.dtype = np.dtype(T)
        self.D = D = len(xs)
        if D not in (2, 3):
            raise ValueError("only 2-D and 3-D setups are supported")
        self.boundary_conditions = dict(boundary_conditions)
        bcu = self.boundary_conditions["u"]
        self.bc_u = bcu
        self.bc_temp = self.boundary_conditions.get("temp")
        self.xlims = tuple((v.min(), v.max()) for v in xs)
        self.device = device

        # ghost coordinates (boundary_conditions.jl:45-64), left then right (grid.jl:105-109)
        for d in range(D):
            for right in (False, True):
                bc = bcu[d][int(right)]
                v = xs[d]
                if isinstance(bc, PeriodicBC):
                    v = np.append(v, v[-1] + (v[1] - v[0])) if right else np.insert(v, 0, v[0] - (v[-1] - v[-2]))
                elif isinstance(bc, DirichletBC):
                    v = np.append(v, v[-1]) if right else np.insert(v, 0, v[0])
                elif isinstance(bc, SymmetricBC):
                    v = np.append(v, v[-1] + (v[-1] - v[-2])) if right else np.insert(v, 0, v[0] - (v[1] - v[0]))
                elif isinstance(bc, PressureBC):
                    v = np.append(v, v[-1]) if right else np.insert(v, 0, [v[0], v[0]])
                else:
                    raise TypeError(f"unknown boundary condition {bc!r}")
                xs[d] = v
        self.x = tuple(xs)
        self.N = tuple(len(v) - 1 for v in xs)
        N = self.N
        code = [[bcu[d][0].code, bcu[d][1].code] for d in range(D)]
        self.bc_codes = code
        self.Ip = tuple((_noff_p(code[d][0], False), N[d] - _noff_p(code[d][1], True)) for d in range(D))
        self.Np = tuple(hi - lo for lo, hi in self.Ip)
        self.Iu = tuple(
            tuple(
                (_noff_u(code[d][0], False, c == d), N[d] - _noff_u(code[d][1], True, c == d))
                for d in range(D)
            )
            for c in range(D)
        )
        self.Nu = tuple(tuple(hi - lo for lo, hi in self.Iu[c]) for c in range(D))
        self.xp = tuple((v[:-1] + v[1:]) / 2 for v in xs)
        self.xu = tuple(
            tuple(xs[d][1:] if c == d else (xs[d][:-1] + xs[d][1:]) / 2 for d in range(D)) for c in range(D)
        )
        eps = np.finfo(T).eps
        self.Delta = tuple(np.ascontiguousarray(np.maximum(np.diff(v), eps), dtype=T) for v in xs)
        self.DeltaU = tuple(
            np.ascontiguousarray(
                np.maximum(np.append(np.diff(self.xp[d]), self.Delta[d][-1] / 2), eps), dtype=T
            )
            for d in range(D)
        )
        A1, A2 = [], []
        for d in range(D):
            n = N[d]
            w2 = (xs[d][1:n] - self.xp[d][: n - 1]) / self.DeltaU[d][: n - 1]
            A1.append(np.ascontiguousarray(np.concatenate(([1], 1 - w2)), dtype=T))
            A2.append(np.ascontiguousarray(np.concatenate((w2, [1])), dtype=T))
        self.A1, self.A2 = tuple(A1), tuple(A2)
        self._plans = {}
        # z-slab decomposition
        self.rank, self.world, self.nccl_id = int(rank), int(world), nccl_id
        if self.world > 1:
            zt = [type(b).__name__ for b in boundary_conditions["u"][2]] if D == 3 else []
            if D != 3 or "PressureBC" in zt or (zt[0] == "PeriodicBC") != (zt[1] == "PeriodicBC"):
                raise ValueError("world > 1 needs a 3-D grid that is periodic, wall-bounded or symmetric in z")
            if (N[2] - 2) % self.world:
                raise ValueError("the number of z volumes must be divisible by world")
            if nccl_id is None or len(nccl_id) != 128:
                raise ValueError("world > 1 needs the 128-byte nccl_id shared by all ranks")
        self.nzl = (N[2] - 2) // self.world if D == 3 else 1
        self.zoff = self.rank * self.nzl if self.world > 1 else 0
        self.N_local = N if self.world == 1 else (N[0], N[1], self.nzl + 2)

    def local_range(self, rng, d):
        """Half-open global index range ``rng`` of dim ``d`` clipped to this rank's slab and
        shifted to local indices (ghost planes excluded from ownership)."""
        lo, hi = rng
        if self.world == 1 or d != 2:
            return lo, hi
        glo, ghi = max(lo, self.zoff + 1), min(hi, self.zoff + self.nzl + 1)
        return glo - self.zoff, max(ghi, glo) - self.zoff

    # -- helpers -------------------------------------------------------------------------
    def is_periodic(self, d):
        return self.bc_codes[d][0] == _lib.BC_PERIODIC and self.bc_codes[d][1] == _lib.BC_PERIODIC

    def is_wall(self, d):
        return self.bc_codes[d][0] == _lib.BC_DIRICHLET and self.bc_codes[d][1] == _lib.BC_DIRICHLET

    def slices(self, rng):
        """Half-open index ranges -> tuple of slices for a (N1, N2[, N3]) NumPy view."""
        return tuple(slice(lo, hi) for lo, hi in rng)

    def bc_functions(self):
        """[(dim, side, fn)] of the function-valued DirichletBC(u), u(dim, x..., t) (boundary_conditions.jl:7-20)"""
        return [(d, side, self.bc_u[d][side].u) for d in range(self.D) for side in (0, 1)
                if isinstance(self.bc_u[d][side], DirichletBC) and callable(self.bc_u[d][side].u)]

    def evaluate_bc_planes(self, t, dudt=False):
        """{(dim, side, comp): plane} of every function-valued DirichletBC at time t: the reference's
        `u[I, α] .= bcfunc.(α, xI..., t)` over the plane `boundary(β, N, Iu[α], isright)` spanning the full
        (ghosted) extent of the other dimensions (src/boundary_conditions.jl:341-372), or its centred
        time derivative with step sqrt(eps)/2 (:348-354).  Plane element (i1, i2) over the remaining
        dimensions sits at i1 + N[d1] * i2 (what insb200_set_bc_u_plane expects)."""
        D, T = self.D, self.T
        out = {}
        for beta, side, fn in self.bc_functions():
            for a in range(D):
                lo, hi = self.Iu[a][beta]
                i = hi if side else lo - 1
                xs = []
                for g in range(D):
                    c = self.xu[a][g][i:i + 1] if g == beta else self.xu[a][g]
                    shape = [1] * D
                    shape[g] = len(c)
                    xs.append(c.reshape(shape))
                if dudt:
                    h = T(np.sqrt(np.finfo(T).eps) / 2)
                    val = (fn(a, *xs, T(t) + h) - fn(a, *xs, T(t) - h)) / (2 * h)
                else:
                    val = fn(a, *xs, T(t))
                shape = [1 if g == beta else self.N[g] for g in range(D)]
                plane = np.squeeze(np.broadcast_to(np.asarray(val, dtype=T), shape), axis=beta)
                out[beta, side, a] = np.array(plane.T if D == 3 else plane, dtype=T, order="C", copy=True)
        return out

    def bc_constants(self):
        """Dirichlet constants as (u_val[d][side][comp], temp_val[d][side]); function-valued
        BCs are not representable here and raise."""
        D = self.D
        uval = np.zeros((3, 2, 3))
        tval = np.zeros((3, 2))
        for d in range(D):
            for side in (0, 1):
                bc = self.bc_u[d][side]
                if isinstance(bc, DirichletBC) and bc.u is not None:
                    if callable(bc.u):
                        # function-valued: evaluated on the plane coordinates by the host and handed
                        # to the library as planes (api.BcPlanes, insb200_set_bc_u_plane)
                        continue
                    for c in range(D):
                        uval[d, side, c] = float(bc.u[c])
                if self.bc_temp is not None:
                    bt = self.bc_temp[d][side]
                    if isinstance(bt, DirichletBC) and bt.u is not None:
                        if not isinstance(bt.u, numbers.Number):
                            raise NotImplementedError("function-valued temperature BC")
                        tval[d, side] = float(bt.u)
        return uval, tval
