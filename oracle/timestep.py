"""Oracle (test infrastructure): time steppers and the driver loop.

Restates /root/reference/src/time_steppers/methods.jl:56-97,
RKMethods.jl:125-134 (Wray3) and :251-257 (RK44),
step_explicit_runge_kutta.jl:4-57 (``timestep!`` for explicit RK),
step_lmwray3.jl:3-94 (``timestep!`` for LMWray3),
time_stepper_caches.jl:11-24, /root/reference/src/solver.jl:77-149
(``solve_unsteady``) and /root/reference/src/sciml.jl:35-48
(``right_hand_side!``).

A *state* is a dict ``{"u": vector field[, "temp": scalar field]}``.
``force_`` is ``(force, state, t) -> None`` (fills ``force`` in place), like the
reference's ``force!(force, state, t; setup, cache, params...)``.
"""

from __future__ import annotations

from fractions import Fraction

import numpy as np

from . import bc as _bc
from . import operators as ops
from .grid import scalarfield
from .pressure import project_


# ---------------------------------------------------------------------------
# methods (methods.jl:56-97)
# ---------------------------------------------------------------------------
class ExplicitRungeKuttaMethod:
    def __init__(self, A, b, c, r, T=np.float64):
        self.A, self.b, self.c, self.r, self.T = A, b, c, r, np.dtype(T).type


def runge_kutta_method(A, b, c, r, T=np.float64):
    """methods.jl:71-89: convert to T, then SHIFT the tableau."""
    T = np.dtype(T).type
    A = np.array([[float(v) for v in row] for row in A], dtype=np.float64)
    b = np.array([float(v) for v in b], dtype=np.float64)
    c = np.array([float(v) for v in c], dtype=np.float64)
    s = A.shape[0]
    if not (s == A.shape[1] == len(b) == len(c)):
        raise ValueError("A, b, and c must have the same sizes")
    if not np.allclose(np.triu(A), 0):
        raise ValueError("Only explicit Runge-Kutta methods are supported")
    A, b, c = A.astype(T), b.astype(T), c.astype(T)
    A = np.vstack([A[1:, :], b[None, :]])
    c = np.append(c[1:], T(1))
    return ExplicitRungeKuttaMethod(A, b, c, T(r), T)


def RK44(T=np.float64):
    """RKMethods.jl:251-257"""
    h = Fraction(1, 2)
    A = [[0, 0, 0, 0], [h, 0, 0, 0], [0, h, 0, 0], [0, 0, 1, 0]]
    b = [Fraction(1, 6), Fraction(1, 3), Fraction(1, 3), Fraction(1, 6)]
    c = [sum(row) for row in A]
    return runge_kutta_method(A, b, c, 0, T)


def Wray3(T=np.float64):
    """RKMethods.jl:125-134"""
    A = [[Fraction(0)] * 3 for _ in range(3)]
    A[1][0] = Fraction(8, 15)
    A[2][0] = Fraction(8, 15) - Fraction(17, 60)
    A[2][1] = Fraction(5, 12)
    b = [Fraction(8, 15) - Fraction(17, 60), Fraction(0), Fraction(3, 4)]
    c = [Fraction(0), A[1][0], A[2][0] + A[2][1]]
    return runge_kutta_method(A, b, c, 0, T)


def FE11(T=np.float64):
    """Forward Euler (RKMethods.jl): smallest explicit tableau, used in tests."""
    return runge_kutta_method([[0]], [1], [0], 1, T)


def RK33C2(T=np.float64):
    """A 3-stage tableau with a dense second row (general-tableau test case;
    same coefficients as RKMethods.RK33C2)."""
    A = [[0, 0, 0], [Fraction(2, 3), 0, 0], [Fraction(1, 3), Fraction(1, 3), 0]]
    b = [Fraction(1, 4), 0, Fraction(3, 4)]
    c = [0, Fraction(2, 3), Fraction(2, 3)]
    return runge_kutta_method(A, b, c, 0, T)


class LMWray3:
    """methods.jl:91-97; coefficients live in step_lmwray3.jl:60-62."""

    def __init__(self, T=np.float64):
        self.T = np.dtype(T).type


# ---------------------------------------------------------------------------
# caches (time_stepper_caches.jl:11-24)
# ---------------------------------------------------------------------------
def _similar(state):
    return {k: np.zeros_like(v, order="F") for k, v in state.items()}


def get_cache(method, state, setup):
    if isinstance(method, ExplicitRungeKuttaMethod):
        return dict(
            statestart=_similar(state),
            k=[_similar(state) for _ in range(len(method.b))],
            p=scalarfield(setup),
        )
    return dict(statestart=_similar(state), k=_similar(state), p=scalarfield(setup))


def create_stepper(method, setup, psolver, state, t, n=0):
    return dict(setup=setup, psolver=psolver, state=state, t=t, n=n)


# ---------------------------------------------------------------------------
# timestep! for explicit RK (step_explicit_runge_kutta.jl:4-57)
# ---------------------------------------------------------------------------
def _f(dt):
    """A Python float stands for a Julia Float64 literal; NumPy scalars keep their type."""
    return dt if isinstance(dt, np.floating) else np.float64(dt)


def _axpy_typed(u, coef, k):
    """``@. u += Δt * A[i,j] * k`` with Julia's promotion: ``coef`` carries the
    coefficient dtype; the sum is evaluated in promote(T_field, T_coef) and
    rounded once on store."""
    ct = np.result_type(u.dtype, np.asarray(coef).dtype)
    if ct == u.dtype:
        u += u.dtype.type(coef) * k
    else:
        u[...] = (u.astype(ct) + coef * k.astype(ct)).astype(u.dtype)


def timestep_rk_(method, force_, stepper, dt, cache):
    setup, psolver, state, t = stepper["setup"], stepper["psolver"], stepper["state"], stepper["t"]
    A, b, c = method.A, method.b, method.c
    statestart, k, p = cache["statestart"], cache["k"], cache["p"]
    nstage = len(b)
    CT = method.T
    tstart = t
    for key in state:
        statestart[key][...] = state[key]
    for i in range(nstage):
        force_(k[i], state, t)
        for key in state:
            state[key][...] = statestart[key]
        for j in range(i + 1):
            coef = _f(dt) * A[i, j]  # Julia: Δt * A[i, j] in promote(typeof(Δt), T_method)
            for key in state:
                _axpy_typed(state[key], coef, k[j][key])
        t = tstart + c[i] * dt
        _bc.apply_bc_u(state["u"], t, setup)
        project_(state["u"], setup, psolver, p)
        for key in state:
            if key == "u":
                _bc.apply_bc_u(state["u"], t, setup)
            elif key == "temp":
                _bc.apply_bc_temp(state["temp"], t, setup)
            else:
                _bc.apply_bc_p(state[key], t, setup)
    return create_stepper(method, setup, psolver, state, t, stepper["n"] + 1)


# ---------------------------------------------------------------------------
# timestep! for LMWray3 (step_lmwray3.jl:3-94)
# ---------------------------------------------------------------------------
def timestep_lmwray3_(method, force_, stepper, dt, cache):
    setup, psolver, state, t = stepper["setup"], stepper["psolver"], stepper["state"], stepper["t"]
    statestart, p, k = cache["statestart"], cache["p"], cache["k"]
    T = state["u"].dtype.type
    dotemp = "temp" in state
    keys = ["u"] + (["temp"] if dotemp else [])
    tstart = t
    for key in keys:
        statestart[key][...] = state[key]
    a = (T(8 / 15), T(5 / 12), T(3 / 4))
    b = (T(1 / 4), T(0))
    c = (T(8 / 15), T(2 / 3), T(1))
    nstage = 3
    for i in range(nstage):
        force_(k, state, t)
        t = tstart + c[i] * dt
        for key in keys:
            state[key][...] = statestart[key]
        for key in keys:  # axpy!(a[i]*Δt, k, state): BLAS converts the scalar to T
            _axpy_typed(state[key], T(a[i] * _f(dt)), k[key])
        _bc.apply_bc_u(state["u"], t, setup)
        project_(state["u"], setup, psolver, p)
        if i != nstage - 1:
            for key in keys:
                _axpy_typed(statestart[key], T(b[i] * _f(dt)), k[key])
        _bc.apply_bc_u(state["u"], t, setup)
        if dotemp:
            _bc.apply_bc_temp(state["temp"], t, setup)
    t = tstart + dt
    _bc.apply_bc_u(state["u"], t, setup)
    if dotemp:
        _bc.apply_bc_temp(state["temp"], t, setup)
    return create_stepper(method, setup, psolver, state, t, stepper["n"] + 1)


def timestep_(method, force_, stepper, dt, cache):
    if isinstance(method, LMWray3):
        return timestep_lmwray3_(method, force_, stepper, dt, cache)
    return timestep_rk_(method, force_, stepper, dt, cache)


# ---------------------------------------------------------------------------
# forces
# ---------------------------------------------------------------------------
def make_navierstokes(setup, viscosity, bodyforce=None):
    """navierstokes! (solver.jl:2-6); bodyforce: the constant body force of examples/ChannelFlow.jl:172-176"""

    def force_(force, state, t):
        ops.navierstokes_(force, state, t, setup, viscosity)
        if bodyforce is not None:
            for c, b in enumerate(bodyforce):
                force["u"][..., c] += setup.T(b)

    force_.params = dict(viscosity=viscosity)
    return force_


def make_navierstokes_eddy(setup, viscosity, model, C, bodyforce=None):
    """force_eddy! of examples/ChannelFlow.jl:177-182: navierstokes! + eddy_viscosity_closure!
    (+ constant body force)."""
    from . import eddyviscosity as ev

    cache = dict(visc=scalarfield(setup), G=ev.tensorfield(setup))

    def force_(force, state, t):
        ops.navierstokes_(force, state, t, setup, viscosity)
        ev.eddy_viscosity_closure_(model, C, force["u"], state["u"], setup, cache)
        if bodyforce is not None:
            for c, b in enumerate(bodyforce):
                force["u"][..., c] += setup.T(b)

    force_.params = dict(viscosity=viscosity)
    return force_


def make_boussinesq(setup, viscosity, conductivity, gdir, gravity, dodissipation):
    def force_(force, state, t):
        ops.boussinesq_(
            force, state, t, setup, viscosity, conductivity, gdir, gravity, dodissipation
        )

    force_.params = dict(viscosity=viscosity, conductivity=conductivity)
    return force_


# ---------------------------------------------------------------------------
# solve_unsteady (solver.jl:77-149)
# ---------------------------------------------------------------------------
def solve_unsteady(
    setup,
    tlims,
    start,
    force_,
    method=None,
    psolver=None,
    dt=None,
    dt_min=None,
    cfl=None,
    n_adapt_dt=1,
    callback=None,
):
    from .pressure import default_psolver

    T = start["u"].dtype.type
    method = method if method is not None else LMWray3(T)
    psolver = psolver if psolver is not None else default_psolver(setup)
    cfl = T(0.9) if cfl is None else cfl
    tstart, tend = tlims
    state = {k: v.copy(order="F") for k, v in start.items()}
    cache = get_cache(method, state, setup)
    stepper = create_stepper(method, setup, psolver, state, tstart)
    if dt is None:
        while stepper["t"] < tend:
            if stepper["n"] % n_adapt_dt == 0:
                dt = cfl * ops.propose_timestep(
                    stepper["state"], setup, **getattr(force_, "params", {})
                )
                if dt_min is not None:
                    dt = max(dt, dt_min)
                if dt < 1e-10 or dt > 100 or np.isnan(dt):
                    break
            dt_step = min(dt, tend - stepper["t"])
            stepper = timestep_(method, force_, stepper, dt_step, cache)
            if callback:
                callback(stepper)
    else:
        nstep = int(round((tend - tstart) / dt))
        dt = (tend - tstart) / nstep
        for _ in range(nstep):
            stepper = timestep_(method, force_, stepper, dt, cache)
            if callback:
                callback(stepper)
    out = dict(stepper["state"])
    out["t"] = stepper["t"]
    return out


# ---------------------------------------------------------------------------
# right_hand_side! (sciml.jl:35-48)
# ---------------------------------------------------------------------------
def right_hand_side_(dudt, u, setup, psolver, viscosity, t):
    p = scalarfield(setup)
    tmp = u.copy(order="F")
    _bc.apply_bc_u(tmp, t, setup)
    ops.navierstokes_({"u": dudt}, {"u": tmp}, t, setup, viscosity)
    _bc.apply_bc_u(dudt, t, setup, dudt=True)
    project_(dudt, setup, psolver, p)
    return None
