"""Python mirror of the reference's operator / solver API for the hot path, bound to
libinsb200.so.  Function names follow the Julia ones with ``!`` spelled ``_``:

    Setup, scalarfield, vectorfield, apply_bc_u_, apply_bc_p_, apply_bc_temp_,
    divergence_, scalewithvolume_, pressuregradient_, applypressure_, convection_,
    diffusion_, navierstokes_, boussinesq_, psolver_spectral, psolver_transform,
    psolver_tridiagonal, default_psolver, poisson_, project_, RKMethods, LMWray3,
    create_stepper, timestep_, propose_timestep, solve_unsteady, right_hand_side_,
    create_right_hand_side, processor, timelogger, fieldsaver

Fields are torch CUDA tensors whose memory is the reference's column-major ghosted
layout: a scalar has shape ``(N3, N2, N1)`` (``(N2, N1)`` in 2-D) and a vector
``(D, N3, N2, N1)`` -- i.e. ``numpy_F_array.T``.  torch only provides device memory and
streams; all arithmetic happens in the CUDA library.
"""

from __future__ import annotations

import ctypes as C
import weakref
import os
import time
from fractions import Fraction

import numpy as np
import torch

from . import _lib
from ._lib import Insb200Error, ParamsStruct, SetupStruct, check, lib
from .grid import DirichletBC, PeriodicBC, PressureBC, Setup, SymmetricBC  # noqa: F401

_TORCH_DTYPE = {np.float32: torch.float32, np.float64: torch.float64}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


# ------------------------------------------------------------------------------------
# plans
# ------------------------------------------------------------------------------------
class BcPlanes:
    """Function-valued DirichletBC (src/boundary_conditions.jl:341-372): the user function cannot
    cross the C ABI, so the host evaluates it on the boundary-plane coordinates -- `bcfunc(α, x..., t)`
    over the plane `boundary(β, N, Iu[α], isright)`, all of the other dimensions -- into small device
    buffers that every plan of the setup holds pointers to (insb200_set_bc_u_plane).  `refresh(t)`
    re-evaluates them (time-dependent functions: once per stage time); `dudt=True` gives the centred
    time derivative with step sqrt(eps)/2 (:348-354) used by right_hand_side!."""

    def __init__(self, setup):
        self.setup = setup
        self.funcs = setup.bc_functions()
        if self.funcs and setup.world > 1:
            raise NotImplementedError("function-valued DirichletBC on a decomposed grid")
        self.planes = {}
        self.state = None
        if not self.funcs:
            self.time_dependent = False
            return
        a = self._evaluate(0.0, False)
        declared = [setup.bc_u[d][side].time_dependent for d, side, _ in self.funcs]
        if any(v is True for v in declared):
            self.time_dependent = True
        elif all(v is False for v in declared):
            self.time_dependent = False
        else:  # undeclared: probe two more times now, and keep probing at every step start (refresh)
            self.time_dependent = any(not np.array_equal(a[k], b[k]) for t in (0.7853981, 1.9098593)
                                      for b in (self._evaluate(t, False),) for k in a)
        self.probing = not self.time_dependent and any(v is None for v in declared)
        for k, v in a.items():
            self.planes[k] = torch.from_numpy(v).to(_dev(setup))
        self.host = a
        self.state = (0.0, False)

    def _evaluate(self, t, dudt):
        return self.setup.evaluate_bc_planes(t, dudt)

    def attach(self, plan):
        for (beta, side, a), tns in self.planes.items():
            check(lib().insb200_set_bc_u_plane(plan.h, beta, side, a, _ptr(tns)))

    def refresh(self, t, dudt=False):
        if not self.funcs:
            return
        if self.probing and not dudt and float(t) != 0.0:
            # a function not declared time independent is re-evaluated at the current time: as soon as its
            # planes differ from the ones in use it is time dependent from here on (a ramp that starts
            # late, a period the construction-time probes happened to hit)
            now = self._evaluate(float(t), False)
            if any(not np.array_equal(now[k], self.host[k]) for k in now):
                self.time_dependent, self.probing = True, False
        want = (float(t) if (self.time_dependent or dudt) else 0.0, bool(dudt))
        if self.state == want:
            return
        for k, v in self._evaluate(want[0], dudt).items():
            self.planes[k].copy_(torch.from_numpy(v))
        self.state = want


def _bc_planes(setup) -> BcPlanes:
    if "bcplanes" not in setup._plans:
        setup._plans["bcplanes"] = BcPlanes(setup)
    return setup._plans["bcplanes"]


class Plan:
    """Owns an ``insb200_plan_t`` (device tables, Poisson workspace, stepper caches)."""

    def __init__(self, setup: Setup, psolver=_lib.PSOLVER_NONE, tridiagonal_dir=-1, eig=None, cg=None):
        if not torch.cuda.is_available():
            raise Insb200Error("no CUDA device: insb200 has no CPU fallback")
        self.setup = setup
        s = SetupStruct()
        s.D = setup.D
        for d in range(3):
            s.N[d] = setup.N[d] if d < setup.D else 1
        s.dtype = _lib.F64 if setup.T is np.float64 else _lib.F32
        uval, tval = setup.bc_constants()
        self._keep = []
        for d in range(setup.D):
            for side in (0, 1):
                s.bc_u[d][side] = setup.bc_codes[d][side]
                if setup.bc_temp is not None:
                    s.bc_temp[d][side] = setup.bc_temp[d][side].code
                for c in range(3):
                    s.bc_u_value[d][side][c] = uval[d, side, c]
                s.bc_temp_value[d][side] = tval[d, side]
                s.xlims[d][side] = float(setup.xlims[d][side])
            for name in ("Delta", "DeltaU", "A1", "A2"):
                arr = np.ascontiguousarray(getattr(setup, name)[d], dtype=setup.T)
                self._keep.append(arr)
                getattr(s, name)[d] = arr.ctypes.data
            if eig is not None and eig[d] is not None:
                arr = np.ascontiguousarray(eig[d], dtype=setup.T)
                self._keep.append(arr)
                s.eig[d] = arr.ctypes.data
        s.has_temp = int(setup.bc_temp is not None)
        s.psolver = psolver
        s.tridiagonal_dir = tridiagonal_dir
        if cg:
            s.cg_abstol, s.cg_reltol = float(cg.get("abstol") or 0.0), float(cg.get("reltol") or 0.0)
            s.cg_maxiter = int(cg.get("maxiter") or 0)
        s.device = torch.cuda.current_device() if setup.device is None else int(setup.device)
        s.rank, s.world = setup.rank, setup.world
        if setup.world > 1:
            idbuf = C.create_string_buffer(bytes(setup.nccl_id), 128)
            self._keep.append(idbuf)
            s.nccl_id = C.cast(idbuf, C.c_void_p)
        h = C.c_void_p()
        check(lib().insb200_plan_create(C.byref(s), C.byref(h)))
        self.h = h
        self.psolver = lib().insb200_plan_psolver(h)
        _bc_planes(setup).attach(self)
        self.p2p = False
        if setup.world > 1 and self.psolver != _lib.PSOLVER_NONE and not os.environ.get("INSB200_NO_P2P"):
            self._enable_p2p()
        setup._plans.setdefault("_registry", weakref.WeakSet()).add(self)  # (release_plans)

    def _enable_p2p(self):
        """Direct peer-store FFT transposes: gather everybody's CUDA IPC handles through the
        torch.distributed group (insb200_plan_ipc_export / _import)."""
        import torch.distributed as dist

        blob = C.create_string_buffer(192)
        check(lib().insb200_plan_ipc_export(self.h, C.cast(blob, C.c_void_p)))
        blobs = [None] * self.setup.world
        dist.all_gather_object(blobs, blob.raw)
        allb = C.create_string_buffer(b"".join(blobs), 192 * self.setup.world)
        check(lib().insb200_plan_ipc_import(self.h, C.cast(allb, C.c_void_p)))
        self.p2p = True

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().insb200_plan_destroy(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def workspace_bytes(self):
        return lib().insb200_plan_workspace_bytes(self.h)

    @property
    def launch_count(self):
        return lib().insb200_launch_count(self.h)

    def profile_enable(self, on=True):
        check(lib().insb200_profile_enable(self.h, int(on)))

    def profile_read(self):
        ms = (C.c_double * 12)()
        cnt = (C.c_longlong * 12)()
        check(lib().insb200_profile_read_ex(self.h, 12, ms, cnt))
        names = ("rhs_combine", "divergence", "poisson", "project", "bc", "fft_x_r2c", "fft_y_fwd", "transpose_fwd",
                 "fft_z_mid", "transpose_bwd", "fft_y_inv", "fft_x_c2r")
        return {n: (ms[i], cnt[i]) for i, n in enumerate(names)}


def release_plans(setup: Setup):
    """Collective tear-down of every plan of a decomposed setup (call on all ranks at the same point): unmap the
    peers' buffers, barrier, destroy, barrier -- an exporting process must not free memory a peer still has mapped,
    and the next setup's IPC handles must not meet stale mappings.  A no-op chain for world == 1."""
    plans = list(setup._plans.get("_registry", ()))
    if setup.world > 1:
        import torch.distributed as dist

        torch.cuda.synchronize()
        for p in plans:
            if p.h:
                check(lib().insb200_plan_ipc_close(p.h))
        dist.barrier()
    for p in plans:
        if p.h:
            lib().insb200_plan_destroy(p.h)
            p.h = None
    setup._plans.clear()
    if setup.world > 1:
        import torch.distributed as dist

        dist.barrier()


def _base_plan(setup: Setup) -> Plan:
    if "base" not in setup._plans:
        setup._plans["base"] = Plan(setup, _lib.PSOLVER_NONE)
    return setup._plans["base"]


class Smagorinsky:
    """Eddy-viscosity model structs of src/eddyviscosity.jl:252-263."""
    code = 1

    def __init__(self, C):
        self.C = float(C)


class WALE(Smagorinsky):
    code = 2


class Vreman(Smagorinsky):
    code = 3


class QR(Smagorinsky):
    code = 4


def _params(setup, viscosity=0.0, conductivity=0.0, gravity=0.0, gdir=1, dodissipation=False,
            boussinesq=False, bodyforce=None, eddyviscosity=None):
    p = ParamsStruct()
    p.eddy_model = 0 if eddyviscosity is None else int(eddyviscosity.code)
    p.eddy_C = 0.0 if eddyviscosity is None else float(eddyviscosity.C)
    p.viscosity = float(viscosity)
    p.conductivity = float(conductivity)
    p.gravity = float(gravity)
    p.gdir = int(gdir)
    p.dodissipation = int(bool(dodissipation))
    p.boussinesq = int(bool(boussinesq))
    for c in range(3):
        p.bodyforce[c] = float(bodyforce[c]) if bodyforce is not None and c < len(bodyforce) else 0.0
    return p


# ------------------------------------------------------------------------------------
# fields (src/initializers.jl:2,5)
# ------------------------------------------------------------------------------------
def _dev(setup):
    return torch.device("cuda", torch.cuda.current_device() if setup.device is None else int(setup.device))


def scalarfield(setup):
    return torch.zeros(tuple(reversed(setup.N_local)), dtype=_TORCH_DTYPE[setup.T], device=_dev(setup))


def vectorfield(setup):
    return torch.zeros((setup.D,) + tuple(reversed(setup.N_local)), dtype=_TORCH_DTYPE[setup.T], device=_dev(setup))


def local_slab(setup, arr):
    """This rank's planes (ghosts included) of a GLOBAL array in the reference layout
    ((N1,N2,N3[,D])); the identity for world == 1."""
    if setup.world == 1:
        return arr
    return arr[:, :, setup.zoff: setup.zoff + setup.nzl + 2]


def to_device(setup, arr):
    """NumPy array in the reference layout ((N1,N2[,N3][,D]), any order) -> torch CUDA tensor.
    A GLOBAL array is cut down to this rank's slab when world > 1."""
    if setup.world > 1 and arr.shape[2] == setup.N[2]:
        arr = local_slab(setup, arr)
    a = np.asfortranarray(arr, dtype=setup.T)
    return torch.from_numpy(np.ascontiguousarray(a.T)).to(_dev(setup))


def comm_unique_id():
    """128-byte NCCL unique id (insb200_comm_unique_id); create on rank 0, share with all ranks."""
    buf = C.create_string_buffer(128)
    check(lib().insb200_comm_unique_id(C.cast(buf, C.c_void_p)))
    return buf.raw


def share_unique_id():
    """Rank 0 of the initialised torch.distributed group creates the id, everybody gets it."""
    import torch.distributed as dist

    obj = [comm_unique_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    return obj[0]


def gather_global(setup, t):
    """All ranks: assemble the GLOBAL ghosted array (reference layout, NumPy) from the slabs.
    Interior planes come from their owners; the two global ghost planes from the end ranks."""
    if setup.world == 1:
        return to_host(t)
    import torch.distributed as dist

    loc = to_host(t)  # (N1, N2, nzl+2[, D])
    parts = [None] * setup.world
    dist.all_gather_object(parts, loc)
    inner = [p[:, :, 1:-1] for p in parts]
    return np.concatenate([parts[0][:, :, :1]] + inner + [parts[-1][:, :, -1:]], axis=2)


def to_host(t):
    """torch tensor -> Fortran-ordered NumPy array in the reference layout."""
    return t.detach().cpu().numpy().T


def _chk(setup, t, vector):
    want = ((setup.D,) if vector else ()) + tuple(reversed(setup.N_local))
    if tuple(t.shape) != want or not t.is_contiguous() or not t.is_cuda:
        raise ValueError(f"expected a contiguous CUDA tensor of shape {want}, got {tuple(t.shape)}")
    if t.dtype != _TORCH_DTYPE[setup.T]:
        raise ValueError("dtype mismatch with the setup")
    return t


# ------------------------------------------------------------------------------------
# boundary conditions and operators
# ------------------------------------------------------------------------------------
def apply_bc_u_(u, t, setup, dudt=False, plan=None):
    _bc_planes(setup).refresh(t, dudt)
    pl = plan or _base_plan(setup)
    check(lib().insb200_apply_bc_u(pl.h, _ptr(_chk(setup, u, True)), float(t), int(dudt), _stream()))
    return u


def apply_bc_p_(p, t, setup, plan=None):
    pl = plan or _base_plan(setup)
    check(lib().insb200_apply_bc_p(pl.h, _ptr(_chk(setup, p, False)), _stream()))
    return p


def apply_bc_temp_(temp, t, setup, plan=None):
    pl = plan or _base_plan(setup)
    check(lib().insb200_apply_bc_temp(pl.h, _ptr(_chk(setup, temp, False)), float(t), _stream()))
    return temp


def apply_bc_u(u, t, setup, dudt=False):
    """apply_bc_u (src/boundary_conditions.jl:148-160): the allocating variant, the argument is left alone."""
    return apply_bc_u_(u.clone(), t, setup, dudt)


def apply_bc_p(p, t, setup):
    return apply_bc_p_(p.clone(), t, setup)


def apply_bc_temp(temp, t, setup):
    return apply_bc_temp_(temp.clone(), t, setup)


def divergence_(div, u, setup):
    check(lib().insb200_divergence(_base_plan(setup).h, _ptr(_chk(setup, div, False)), _ptr(_chk(setup, u, True)), _stream()))
    return div


def divergence(u, setup):
    return divergence_(scalarfield(setup), u, setup)


def scalewithvolume_(p, setup):
    check(lib().insb200_scalewithvolume(_base_plan(setup).h, _ptr(_chk(setup, p, False)), _stream()))
    return p


def scalewithvolume(p, setup):
    """scalewithvolume (src/operators.jl:203-216): allocating variant."""
    return scalewithvolume_(p.clone(), setup)


def pressuregradient_(G, p, setup):
    check(lib().insb200_pressuregradient(_base_plan(setup).h, _ptr(_chk(setup, G, True)), _ptr(_chk(setup, p, False)), _stream()))
    return G


def pressuregradient(p, setup):
    return pressuregradient_(vectorfield(setup), p, setup)


def applypressure_(u, p, setup):
    check(lib().insb200_applypressure(_base_plan(setup).h, _ptr(_chk(setup, u, True)), _ptr(_chk(setup, p, False)), _stream()))
    return u


def convection_(f, u, setup):
    check(lib().insb200_convection(_base_plan(setup).h, _ptr(_chk(setup, f, True)), _ptr(_chk(setup, u, True)), _stream()))
    return f


def diffusion_(f, u, setup, viscosity):
    check(lib().insb200_diffusion(_base_plan(setup).h, _ptr(_chk(setup, f, True)), _ptr(_chk(setup, u, True)), float(viscosity), _stream()))
    return f


def convection(u, setup):
    return convection_(torch.zeros_like(u), u, setup)


def diffusion(u, setup, viscosity):
    return diffusion_(torch.zeros_like(u), u, setup, viscosity)


def navierstokes_(force, state, t, *, setup, cache=None, viscosity, bodyforce=None, eddyviscosity=None):
    """navierstokes! (src/solver.jl:2-6): force["u"] is overwritten (with `eddyviscosity` the
    eddy_viscosity_closure! term is added, examples/ChannelFlow.jl:177-182)."""
    pr = _params(setup, viscosity=viscosity, bodyforce=bodyforce, eddyviscosity=eddyviscosity)
    check(lib().insb200_rhs(_base_plan(setup).h, _ptr(_chk(setup, state["u"], True)), None,
                            _ptr(_chk(setup, force["u"], True)), None, C.byref(pr), _stream()))
    return force


def boussinesq_(force, state, t, *, setup, cache=None, viscosity, conductivity, gdir, gravity, dodissipation):
    """boussinesq! (src/solver.jl:16-35)."""
    pr = _params(setup, viscosity, conductivity, gravity, gdir, dodissipation, True)
    check(lib().insb200_rhs(_base_plan(setup).h, _ptr(_chk(setup, state["u"], True)),
                            _ptr(_chk(setup, state["temp"], False)), _ptr(_chk(setup, force["u"], True)),
                            _ptr(_chk(setup, force["temp"], False)), C.byref(pr), _stream()))
    return force


def navierstokes(state, t, *, setup, viscosity):
    """navierstokes (src/solver.jl:9-13), the allocating right-hand side: c = convection(u), d = diffusion(u), c + d
    -- summed after the two operators, so it differs from navierstokes! by rounding like in the reference
    (test/operators.jl:152-167); diffusion! adds into the convection result (0 + c is exact)."""
    f = convection(state["u"], setup)
    return {"u": diffusion_(f, state["u"], setup, viscosity)}


def boussinesq(state, t, *, setup, viscosity, conductivity, gdir, gravity, dodissipation):
    """boussinesq (src/solver.jl:38-48), allocating; evaluated by the fused right-hand side (insb200_rhs), i.e. with
    the rounding of boussinesq! rather than of the operator-by-operator sum."""
    force = {"u": torch.zeros_like(state["u"]), "temp": torch.zeros_like(state["temp"])}
    return boussinesq_(force, state, t, setup=setup, viscosity=viscosity, conductivity=conductivity, gdir=gdir,
                       gravity=gravity, dodissipation=dodissipation)


def get_cache(obj, *args, **kwargs):
    """get_cache (src/solver.jl:55-61, time_steppers/time_stepper_caches.jl:11-24): the built-in forces have none, and
    the stage arrays of the time steppers are owned by the plan (allocated on first use), so there is nothing for the
    caller to hold; accepted for signature parity."""
    return None


# ------------------------------------------------------------------------------------
# pressure solvers (src/pressure.jl)
# ------------------------------------------------------------------------------------
class PSolver:
    """Callable like the reference closures: ``psolver(p)`` solves in place on the ghosted
    array and returns it.  Owns the plan that the fused time steppers use."""

    def __init__(self, setup, kind, **kw):
        self.setup = setup
        self.plan = Plan(setup, kind, **kw)
        self.kind = {0: "spectral", 1: "transform", 2: "tridiagonal", 4: "cg", -1: "none"}[self.plan.psolver]
        if self.plan.psolver == _lib.PSOLVER_NONE:
            raise Insb200Error("no Poisson solver is available for this setup (decomposed plans need an FFT-type solver)")

    @property
    def iterations(self):
        """iterations of the last psolver_cg solve (0 for the FFT-type solvers)"""
        return int(lib().insb200_psolver_iterations(self.plan.h))

    def __call__(self, p):
        check(lib().insb200_psolve(self.plan.h, _ptr(_chk(self.setup, p, False)), _stream()))
        return p


def psolver_spectral(setup, eig=None):
    return PSolver(setup, _lib.PSOLVER_SPECTRAL, eig=eig)


def psolver_transform(setup, eig=None):
    return PSolver(setup, _lib.PSOLVER_TRANSFORM, eig=eig)


def psolver_tridiagonal(setup, dir=None, eig=None):
    return PSolver(setup, _lib.PSOLVER_TRIDIAGONAL, tridiagonal_dir=-1 if dir is None else int(dir), eig=eig)


def psolver_cg(setup, abstol=None, reltol=None, maxiter=None):
    """psolver_cg (src/pressure.jl:155-203): preconditioned conjugate gradients, any BC / grid."""
    return PSolver(setup, _lib.PSOLVER_CG, cg=dict(abstol=abstol, reltol=reltol, maxiter=maxiter))


def default_psolver(setup):
    """default_psolver (src/pressure.jl:58-78); where the reference falls back to the sparse direct
    solver this returns psolver_cg."""
    return PSolver(setup, _lib.PSOLVER_DEFAULT)


def poisson_(psolver, f):
    return psolver(f)


def poisson(psolver, f):
    """poisson (src/pressure.jl:16): poisson!(psolver, copy(f))."""
    return psolver(f.clone())


def project_(u, setup, *, psolver, p):
    check(lib().insb200_project(psolver.plan.h, _ptr(_chk(setup, u, True)), _ptr(_chk(setup, p, False)), _stream()))
    return u


def project(u, setup, *, psolver):
    return project_(u.clone(), setup, psolver=psolver, p=scalarfield(setup))


# ------------------------------------------------------------------------------------
# time stepping (src/time_steppers/*.jl)
# ------------------------------------------------------------------------------------
class ExplicitRungeKuttaMethod:
    def __init__(self, A, b, c, r, T=np.float64):
        self.A, self.b, self.c, self.r, self.T = A, b, c, r, np.dtype(T).type


def runge_kutta_method(A, b, c, r, T=np.float64):
    """methods.jl:71-89: convert to T, then shift the tableau."""
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
    return ExplicitRungeKuttaMethod(np.vstack([A[1:, :], b[None, :]]), b, np.append(c[1:], T(1)), T(r), T)


class RKMethods:
    """The tableaus of src/time_steppers/RKMethods.jl used on the hot path (any explicit
    tableau can be passed through ``runge_kutta_method``)."""

    @staticmethod
    def FE11(T=np.float64):
        return runge_kutta_method([[0]], [1], [0], 1, T)

    @staticmethod
    def RK44(T=np.float64):
        h = Fraction(1, 2)
        A = [[0, 0, 0, 0], [h, 0, 0, 0], [0, h, 0, 0], [0, 0, 1, 0]]
        return runge_kutta_method(A, [Fraction(1, 6), Fraction(1, 3), Fraction(1, 3), Fraction(1, 6)],
                                  [sum(r) for r in A], 0, T)

    @staticmethod
    def Wray3(T=np.float64):
        A = [[Fraction(0)] * 3 for _ in range(3)]
        A[1][0] = Fraction(8, 15)
        A[2][0] = Fraction(8, 15) - Fraction(17, 60)
        A[2][1] = Fraction(5, 12)
        b = [Fraction(8, 15) - Fraction(17, 60), Fraction(0), Fraction(3, 4)]
        return runge_kutta_method(A, b, [Fraction(0), A[1][0], A[2][0] + A[2][1]], 0, T)

    @staticmethod
    def RK33C2(T=np.float64):
        A = [[0, 0, 0], [Fraction(2, 3), 0, 0], [Fraction(1, 3), Fraction(1, 3), 0]]
        return runge_kutta_method(A, [Fraction(1, 4), 0, Fraction(3, 4)], [0, Fraction(2, 3), Fraction(2, 3)], 0, T)


# The other literal explicit tableaus of src/time_steppers/RKMethods.jl (the families built by linear solves -- SSP104,
# rSSPs2, rSSPs3 -- go through runge_kutta_method directly): name -> (A rows "a b; c d", b, c or None = row sums, r).
_TABLEAUS = {
    "DOPRI6": ("0 0 0 0 0 0; 1/5 0 0 0 0 0; 3/40 9/40 0 0 0 0; 44/45 -56/15 32/9 0 0 0; 19372/6561 -25360/2187 64448/6561 -212/729 0 0; 9017/3168 -355/33 46732/5247 49/176 -5103/18656 0",
              "35/384 0 500/1113 125/192 -2187/6784 11/84", None, "0"),
    "HEM3": ("0 0 0; 1/3 0 0; -1 2 0",
            "0 3/4 1/4", None, "0"),
    "HEM3BS": ("0 0 0; 1/2 0 0; -1 2 0",
              "1/6 2/3 1/6", None, "0"),
    "Heun33": ("0 0 0; 1/3 0 0; 0 2/3 0",
              "1/4 0 3/4", None, "0"),
    "MTE22": ("0 0; 2/3 0",
             "1/4 3/4", None, "1/2"),
    "Mid22": ("0 0; 1/2 0",
             "0 1", None, "1/2"),
    "NSSP21": ("0 0; 3/4 0",
              "0 1", None, "0"),
    "NSSP32": ("0 0 0; 1/3 0 0; 0 1 0",
              "1/2 0 1/2", None, "0"),
    "NSSP33": ("0 0 0; -4/9 0 0; 7/6 -1/2 0",
              "1/4 0 3/4", None, "0"),
    "NSSP53": ("0 0 0 0 0; 1/7 0 0 0 0; 0 3/16 0 0 0; 0 0 1/3 0 0; 0 0 0 2/3 0",
              "1/4 0 0 0 3/4", None, "0"),
    "RK33P2": ("0 0 0; 1/3 0 0; -1 2 0",
              "0 3/4 1/4", None, "0"),
    "RK44C2": ("0 0 0 0; 1/4 0 0 0; 0 1/2 0 0; 1 -2 2 0",
              "1/6 0 2/3 1/6", None, "0"),
    "RK44C23": ("0 0 0 0; 1/2 0 0 0; 1/4 1/4 0 0; 0 -1 2 0",
               "1/6 0 2/3 1/6", None, "0"),
    "RK44P2": ("0 0 0 0; 1 0 0 0; 3/8 1/8 0 0; -1/8 -3/8 3/2 0",
              "1/6 -1/18 2/3 2/9", None, "0"),
    "RK56": ("0 0 0 0 0 0; 1/4 0 0 0 0 0; 1/8 1/8 0 0 0 0; 0 0 1/2 0 0 0; 3/16 -3/8 3/8 9/16 0 0; -3/7 8/7 6/7 -12/7 8/7 0",
            "7/90 0 16/45 2/15 16/45 7/90", None, "0"),
    "SSP22": ("0 0; 1 0",
             "1/2 1/2", None, "1"),
    "SSP33": ("0 0 0; 1 0 0; 1/4 1/4 0",
             "1/6 1/6 2/3", None, "1"),
    "SSP42": ("0 0 0 0; 1/3 0 0 0; 1/3 1/3 0 0; 1/3 1/3 1/3 0",
             "1/4 1/4 1/4 1/4", None, "3"),
    "SSP43": ("0 0 0 0; 1/2 0 0 0; 1/2 1/2 0 0; 1/6 1/6 1/6 0",
             "1/6 1/6 1/6 1/2", None, "2"),
}


def _rat(txt):
    return [Fraction(v) for v in txt.split()]


def _tableau_method(name):
    def make(T=np.float64):
        A, b, c, r = _TABLEAUS[name]
        A = [_rat(row) for row in A.split(";")]
        return runge_kutta_method(A, _rat(b), [sum(row) for row in A] if c is None else _rat(c), Fraction(r), T)

    make.__name__ = name
    make.__doc__ = f"RKMethods.{name} (src/time_steppers/RKMethods.jl)"
    return staticmethod(make)


for _name in _TABLEAUS:
    setattr(RKMethods, _name, _tableau_method(_name))


def _hem5(T=np.float64):
    """RKMethods.HEM5 (Brasey and Hairer, 5 stages, 4th order): irrational entries, evaluated in Float64 first."""
    q = np.sqrt(6.0)
    A = [[0, 0, 0, 0, 0], [3 / 10, 0, 0, 0, 0], [(1 + q) / 30, (11 - 4 * q) / 30, 0, 0, 0],
         [(-79 - 31 * q) / 150, (-1 - 4 * q) / 30, (24 + 11 * q) / 25, 0, 0],
         [(14 + 5 * q) / 6, (-8 + 7 * q) / 6, (-9 - 7 * q) / 4, (9 - q) / 4, 0]]
    return runge_kutta_method(A, [0, 0, (16 - q) / 36, (16 + q) / 36, 1 / 9], [sum(r) for r in A], 0, T)


RKMethods.HEM5 = staticmethod(_hem5)


class LMWray3:
    def __init__(self, T=np.float64):
        self.T = np.dtype(T).type


def create_stepper(method, *, setup, psolver, state, t, n=0):
    return dict(setup=setup, psolver=psolver, state=state, t=t, n=n)


def get_state(stepper):
    out = dict(stepper["state"])
    out["t"], out["n"] = stepper["t"], stepper["n"]
    return out


def _force_params(force, setup, params):
    params = dict(params or {})
    if force is navierstokes_:
        return _params(setup, viscosity=params["viscosity"], bodyforce=params.get("bodyforce"),
                       eddyviscosity=params.get("eddyviscosity"))
    if force is boussinesq_:
        return _params(setup, params["viscosity"], params["conductivity"], params["gravity"], params["gdir"],
                       params["dodissipation"], True, params.get("bodyforce"))
    raise NotImplementedError("only the built-in navierstokes_/boussinesq_ forces run fused on the B200 path")


def _coef_dtype(method_T, dt):
    dt_T = np.dtype(type(dt)).type if isinstance(dt, np.floating) else np.float64
    return _lib.F64 if np.result_type(method_T, dt_T) == np.float64 else _lib.F32


def _stage_cache(stepper, nstage):
    """ode_cache of the host-composed steppers (get_cache, src/time_steppers/*.jl): statestart, the stage
    arrays k and the pressure scratch p, allocated once per (setup, state keys) and kept on the setup."""
    setup, state = stepper["setup"], stepper["state"]
    keys = tuple(k for k in ("u", "temp") if k in state)
    cache = setup._plans.setdefault("stage_cache", {})
    c = cache.get(keys)
    if c is None or len(c["k"]) < nstage:
        mk = {"u": vectorfield, "temp": scalarfield}
        c = {"start": {kk: mk[kk](setup) for kk in keys}, "p": scalarfield(setup),
             "k": [{kk: mk[kk](setup) for kk in keys} for _ in range(nstage)]}
        cache[keys] = c
    return keys, c


def _timestep_staged(method, force, stepper, dt, params):
    """timestep! composed stage by stage on the host (step_explicit_runge_kutta.jl:4-57,
    step_lmwray3.jl:3-94): user-defined force functions, and setups whose Dirichlet functions depend on
    time -- their boundary planes are re-evaluated at every stage time tstart + c[i] dt before apply_bc_u!.
    Every device operation is the library's: the force (insb200_rhs for the built-in ones), the stage
    combine + ghost fills + projection (insb200_stage_combine_project), the statestart update (insb200_axpy)."""
    setup, psolver, state, t0 = stepper["setup"], stepper["psolver"], stepper["state"], stepper["t"]
    T = setup.T
    wray = isinstance(method, LMWray3)
    nstage = 1 if wray else len(method.b)
    keys, cache = _stage_cache(stepper, nstage)
    start, p, ks = cache["start"], cache["p"], cache["k"]
    h, st = psolver.plan.h, _stream()
    planes = _bc_planes(setup)
    ncomp = {"u": setup.D, "temp": 1}
    for kk in keys:  # copyto!(statestart, state)
        check(lib().insb200_stage_combine(h, _ptr(start[kk]), _ptr(_chk(setup, state[kk], kk == "u")), 0, None, None,
                                          _lib.F64, ncomp[kk], st))
    temp = state.get("temp")

    def combine_project(nk, coefs, t):
        planes.refresh(t)
        cf = (C.c_double * max(nk, 1))(*coefs)
        ku = (C.c_void_p * max(nk, 1))(*[_ptr(ks[j]["u"]) for j in range(nk)])
        kt = (C.c_void_p * max(nk, 1))(*[_ptr(ks[j]["temp"]) for j in range(nk)]) if temp is not None else None
        check(lib().insb200_stage_combine_project(
            h, _ptr(state["u"]), _ptr(start["u"]), _ptr(temp), _ptr(start["temp"]) if temp is not None else None, nk,
            ku, kt, cf, cdt, _ptr(p), float(t), st))

    if wray:
        cdt = _lib.F64 if T is np.float64 else _lib.F32  # axpy! converts the scalar to the field type (:60-62)
        a3, b3, c3 = (T(8 / 15), T(5 / 12), T(3 / 4)), (T(1 / 4), T(0)), (T(8 / 15), T(2 / 3), T(1))
        for i in range(3):
            force(ks[0], state, t0 + (0 if i == 0 else float(c3[i - 1]) * dt), setup=setup, **params)
            combine_project(1, [float(T(a3[i] * T(dt)))], t0 + float(c3[i]) * dt)
            if i != 2:
                for kk in keys:
                    check(lib().insb200_axpy(h, _ptr(start[kk]), float(T(b3[i] * T(dt))), _ptr(ks[0][kk]), cdt,
                                             ncomp[kk], st))
    else:
        cdt = _coef_dtype(method.T, dt)
        A, c = np.asarray(method.A), method.c

        def coef(i, j):  # dt * A[i, j] in the coefficient type (include/insb200.h)
            if cdt == _lib.F32:
                return float(np.float32(dt) * np.float32(A[i, j]))
            return float(dt) * float(A[i, j])

        t = t0
        for i in range(nstage):
            force(ks[i], state, t, setup=setup, **params)
            t = t0 + float(c[i]) * dt
            combine_project(i + 1, [coef(i, j) for j in range(i + 1)], t)
    return create_stepper(method, setup=setup, psolver=psolver, state=state, t=t0 + dt, n=stepper["n"] + 1)


def timestep_(method, force, stepper, dt, *, params=None, ode_cache=None, force_cache=None):
    """timestep! -- one fused library call per step."""
    setup, psolver, state = stepper["setup"], stepper["psolver"], stepper["state"]
    planes = _bc_planes(setup)
    planes.refresh(stepper["t"])  # (also probes undeclared functions for time dependence)
    if planes.time_dependent or (force is not navierstokes_ and force is not boussinesq_):
        # Dirichlet functions of time, or a user-defined force (src/solver.jl:55-61): stage-by-stage composition
        return _timestep_staged(method, force, stepper, dt, dict(params or {}))
    pr = _force_params(force, setup, params)
    u = _chk(setup, state["u"], True)
    temp = _chk(setup, state["temp"], False) if "temp" in state and pr.boussinesq else None
    t = stepper["t"]
    if isinstance(method, LMWray3):
        check(lib().insb200_step_lmwray3(psolver.plan.h, _ptr(u), _ptr(temp), float(t), float(dt), C.byref(pr), _stream()))
    else:
        s = len(method.b)
        A = (C.c_double * (s * s))(*[float(v) for v in np.asarray(method.A).reshape(-1)])
        c = (C.c_double * s)(*[float(v) for v in method.c])
        check(lib().insb200_step_rk(psolver.plan.h, A, c, s, _coef_dtype(method.T, dt), _ptr(u), _ptr(temp),
                                    float(t), float(dt), C.byref(pr), _stream()))
    return create_stepper(method, setup=setup, psolver=psolver, state=state, t=t + dt, n=stepper["n"] + 1)


def timestep(method, force, stepper, dt, *, params=None):
    """timestep (src/time_steppers/step.jl:17, step_explicit_runge_kutta.jl:59-101): the non-mutating variant -- the
    state of `stepper` is left alone, the new stepper holds fresh arrays."""
    fresh = dict(stepper)
    fresh["state"] = {k: v.clone() for k, v in stepper["state"].items()}
    return timestep_(method, force, fresh, dt, params=params)


def propose_timestep(force, state, setup, params, plan=None):
    pr = _force_params(force, setup, params)
    out = C.c_double()
    pl = plan or _base_plan(setup)
    check(lib().insb200_cfl(pl.h, _ptr(_chk(setup, state["u"], True)), C.byref(pr), C.byref(out), _stream()))
    return setup.T(out.value)


def maxabs(u, setup, plan=None):
    out = C.c_double()
    check(lib().insb200_maxabs((plan or _base_plan(setup)).h, _ptr(u), C.byref(out), _stream()))
    return out.value


def eddy_viscosity_closure_(eddyvisc, f, u, cache, setup):
    """eddy_viscosity_closure! (src/eddyviscosity.jl:356-378): f -= div(-2 nu_t S); the cache is
    plan-owned scratch (`cache` is accepted for signature parity and ignored)."""
    check(lib().insb200_eddy_viscosity_closure(_base_plan(setup).h, int(eddyvisc.code), float(eddyvisc.C),
                                               _ptr(_chk(setup, f, True)), _ptr(_chk(setup, u, True)), _stream()))
    return None


# ------------------------------------------------------------------------------------
# diagnostics (src/operators.jl:853-1066)
# ------------------------------------------------------------------------------------
def vorticity_(omega, u, setup):
    check(lib().insb200_vorticity(_base_plan(setup).h, _ptr(_chk(setup, omega, setup.D == 3)), _ptr(_chk(setup, u, True)),
                                  _stream()))
    return omega


def vorticity(u, setup):
    return vorticity_(scalarfield(setup) if setup.D == 2 else vectorfield(setup), u, setup)


def interpolate_u_p_(up, u, setup):
    check(lib().insb200_interpolate_u_p(_base_plan(setup).h, _ptr(_chk(setup, up, True)), _ptr(_chk(setup, u, True)),
                                        _stream()))
    return up


def interpolate_u_p(u, setup):
    return interpolate_u_p_(vectorfield(setup), u, setup)


def interpolate_omega_p_(omegap, omega, setup):
    v = setup.D == 3
    check(lib().insb200_interpolate_omega_p(_base_plan(setup).h, _ptr(_chk(setup, omegap, v)), _ptr(_chk(setup, omega, v)),
                                            _stream()))
    return omegap


def interpolate_omega_p(omega, setup):
    return interpolate_omega_p_(scalarfield(setup) if setup.D == 2 else vectorfield(setup), omega, setup)


def qcrit_(q, u, setup):
    check(lib().insb200_qcrit(_base_plan(setup).h, _ptr(_chk(setup, q, False)), _ptr(_chk(setup, u, True)), _stream()))
    return q


def qcrit(u, setup):
    return qcrit_(scalarfield(setup), u, setup)


def kinetic_energy_(ke, u, setup, interpolate_first=False):
    check(lib().insb200_kinetic_energy(_base_plan(setup).h, _ptr(_chk(setup, ke, False)), _ptr(_chk(setup, u, True)),
                                       int(interpolate_first), _stream()))
    return ke


def kinetic_energy(u, setup, interpolate_first=False):
    return kinetic_energy_(scalarfield(setup), u, setup, interpolate_first)


def total_kinetic_energy(u, setup, interpolate_first=False, plan=None):
    out = C.c_double()
    check(lib().insb200_total_kinetic_energy((plan or _base_plan(setup)).h, _ptr(_chk(setup, u, True)),
                                             int(interpolate_first), C.byref(out), _stream()))
    return setup.T(out.value)


def spectral_stuff(setup, kmax=None):
    """spectral_stuff (src/spectral.jl:39-67) on the host: kappa = 0..kmax, per shell the (0-based, first index
    fastest) linear indices into the RFFT-sized array of the modes with kappa <= |k| < kappa + 1 (`inds`), the same
    with the modes whose conjugate is not stored repeated (`energyinds`), and Nhat.  The library bins the shells inside
    its own kernel (insb200_energy_spectrum); this is for callers that index spectra themselves."""
    Np, D = setup.Np, setup.D
    kmax = min(Np) // 2 if kmax is None else int(kmax)
    Nhat = (Np[0] // 2 + 1,) + tuple(Np[1:])
    k2 = (np.arange(Nhat[0]) ** 2).reshape((-1,) + (1,) * (D - 1))
    for a in range(1, D):
        ka = np.fft.fftfreq(Np[a], 1.0 / Np[a]).astype(np.int64)
        shape = [1] * D
        shape[a] = -1
        k2 = k2 + (ka**2).reshape(shape)
    k2 = k2.reshape(-1, order="F")
    isort = np.argsort(k2, kind="stable")
    k2s = k2[isort]
    inds, energyinds = [], []
    for kap in range(kmax + 1):
        shell = isort[np.searchsorted(k2s, kap**2, "left"):np.searchsorted(k2s, (kap + 1) ** 2, "left")]
        kx = shell % Nhat[0]
        conj = shell[(kx > 0) & (kx < Nhat[0] - 1)]
        inds.append(shell)
        energyinds.append(np.concatenate([shell, conj]))
    return dict(kappa=np.arange(kmax + 1), inds=inds, energyinds=energyinds, Nhat=Nhat)


def energyspectrum(u, setup, psolver, kmax=None):
    """energyspectrum (src/spectral.jl:75-117) -> dict(kappa, ehat); `psolver` must be the
    psolver_spectral of this setup (its FFT passes and workspace are used)."""
    kmax = min(setup.Np) // 2 if kmax is None else int(kmax)
    out = (C.c_double * (kmax + 1))()
    check(lib().insb200_energy_spectrum(psolver.plan.h, _ptr(_chk(setup, u, True)), kmax, out, _stream()))
    return dict(kappa=np.arange(kmax + 1), ehat=np.array(out[:], dtype=setup.T))


def turbulence_statistics(u, setup, viscosity, psolver):
    """turbulence_statistics (src/spectral.jl:144-257): Kolmogorov scaling in 3-D, Kraichnan-Batchelor-Leith scaling
    in 2-D, from the spectral sums insb200_spectral_moments evaluates behind the psolver_spectral plan's FFT passes
    (uniform periodic grids only, like the reference's assert_uniform_periodic)."""
    if getattr(psolver, "kind", None) != "spectral":
        raise AssertionError("Turbulence statistics requires a uniform periodic grid (psolver_spectral plan)")
    m = (C.c_double * 3)()
    check(lib().insb200_spectral_moments(psolver.plan.h, _ptr(_chk(setup, u, True)), m, _stream()))
    T = setup.T
    nu, e = T(viscosity), T(m[0])
    diss = T(2) * nu * T(m[1])
    kmax = T(np.pi) * min(setup.Np[a] / (setup.xlims[a][1] - setup.xlims[a][0]) for a in range(setup.D))
    if setup.D == 3:
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
    enstrophy_diss = T(2) * nu * T(m[2])
    l_kra = (nu**3 / enstrophy_diss) ** T(1 / 6)
    l_omega = uavg / omega_rms
    return dict(e=e, uavg=uavg, diss=diss, enstrophy=enstrophy, palinstrophy=enstrophy_diss / (2 * nu),
                enstrophy_diss=enstrophy_diss, omega_rms=omega_rms, l_kra=l_kra, l_omega=l_omega, t_eddy=1 / omega_rms,
                t_ens=enstrophy_diss ** (-T(1 / 3)), Re=uavg * l_omega / nu, kmax_etaK=kmax * l_kra)


# ------------------------------------------------------------------------------------
# processors (src/processors.jl): the call-back boundary of solve_unsteady
# ------------------------------------------------------------------------------------
class Observable:
    def __init__(self, value):
        self._v, self._l = value, []

    def on(self, f):
        self._l.append(f)
        return f

    @property
    def value(self):
        return self._v

    @value.setter
    def value(self, v):
        self._v = v
        for f in self._l:
            f(v)


class processor:
    def __init__(self, initialize, finalize=lambda initialized, state: initialized, nupdate=1):
        self.initialize, self.finalize = initialize, finalize
        self.nupdate = int(nupdate)  # the observers only act on steps that are multiples of this


def timelogger(setup, nupdate=1, out=None):
    """processors.jl:45-80: iteration, t, Δt, umax, wall time per iteration."""
    log = [] if out is None else out

    def initialize(state):
        told = [state.value["t"]]
        wold = [time.time()]

        def cb(st):
            if st["n"] % nupdate:
                return
            now = time.time()
            log.append(dict(n=st["n"], t=float(st["t"]), dt=float(st["t"] - told[0]) / nupdate,
                            umax=maxabs(st["u"], setup), itertime=(now - wold[0]) / nupdate))
            told[0], wold[0] = st["t"], now

        state.on(cb)
        return log

    return processor(initialize, nupdate=nupdate)


def fieldsaver(setup, nupdate=1):
    """processors.jl:234-243: host copies of the state every nupdate steps."""

    def initialize(state):
        states = []

        def cb(st):
            if st["n"] % nupdate == 0:
                states.append({k: (to_host(v).copy() if torch.is_tensor(v) else v) for k, v in st.items()})

        state.on(cb)
        cb(state.value)
        return states

    return processor(initialize, nupdate=nupdate)


_STEP_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_longlong, C.c_double)


def _solve_native(setup, method, force, stepper, tend, dt, dt_min, cfl, n_adapt_dt, params, obs, every, graph):
    """The time loop inside the library (insb200_solve_unsteady): one call for the whole run,
    a CUDA-graph replay per step for fixed dt on one GPU, call-backs only every `every` steps."""
    import math
    import warnings

    psolver, state = stepper["psolver"], stepper["state"]
    pr = _force_params(force, setup, params)
    u = _chk(setup, state["u"], True)
    temp = _chk(setup, state["temp"], False) if "temp" in state and pr.boussinesq else None
    tstart = stepper["t"]
    adaptive = dt is None
    tT = setup.T if adaptive else (np.dtype(type(dt)).type if isinstance(dt, np.floating) else np.float64)
    is_wray = isinstance(method, LMWray3)
    s = 0 if is_wray else len(method.b)
    A = None if is_wray else (C.c_double * (s * s))(*[float(v) for v in np.asarray(method.A).reshape(-1)])
    cc = None if is_wray else (C.c_double * s)(*[float(v) for v in method.c])
    coef = _lib.F64 if is_wray else _coef_dtype(method.T, setup.T(0) if adaptive else dt)

    _bc_planes(setup).refresh(tstart)  # planes may hold d/dt values of an earlier right_hand_side! call
    failure = []

    def on_step(_user, n, t):
        # an exception must not vanish inside ctypes: remember it, stop the loop, re-raise below
        try:
            stepper["t"], stepper["n"] = tT(t), int(n)
            obs.value = get_state(stepper)
            return 0
        except BaseException as e:  # noqa: BLE001
            failure.append(e)
            return 1

    cb = _STEP_CB(on_step)
    t_out, n_out = C.c_double(), C.c_longlong()
    rc = lib().insb200_solve_unsteady(
        psolver.plan.h, int(is_wray), A, cc, s, coef, _ptr(u), _ptr(temp), float(tstart), float(tend),
        -1.0 if adaptive else float(dt), -1.0 if dt_min is None else float(dt_min), float(cfl), int(n_adapt_dt),
        _lib.F32 if tT is np.float32 else _lib.F64, C.byref(pr), int(every), C.cast(cb, C.c_void_p), None, int(graph),
        C.byref(t_out), C.byref(n_out), _stream())
    if failure:
        raise failure[0]
    if rc == 2:  # INSB200_STOPPED_BAD_DT: the reference warns and leaves the loop (solver.jl:113-116)
        warnings.warn(lib().insb200_last_error().decode())
    else:
        check(rc)
    stepper["t"], stepper["n"] = tT(t_out.value), int(n_out.value)
    if every == 0 or stepper["n"] % max(every, 1):
        obs.value = get_state(stepper)  # final state for finalize()
    return stepper


def _interior(setup, t, vector):
    """host copy of view(f, Ip[, :]) in the reference layout -- the device -> host staging of a
    plotted / saved frame moves the interior of ONE derived field, not the state"""
    a = to_host(t)
    return np.ascontiguousarray(a[setup.slices(setup.Ip) + ((slice(None),) if vector else ())])


def observefield(state, *, setup, fieldname, psolver=None):
    """observefield (src/processors.jl:85-144): an Observable holding the host copy of a derived
    field on the pressure points; the derived fields are computed on the device (diag.cu)."""
    state = state if isinstance(state, Observable) else Observable(state)
    D = setup.D
    buf = {}

    def compute(st):
        u = st["u"]
        if fieldname in (1, 2, 3):
            buf["up"] = interpolate_u_p_(buf.get("up") if "up" in buf else vectorfield(setup), u, setup)
            return _interior(setup, buf["up"][fieldname - 1], False)
        if fieldname == "velocity":
            buf["up"] = interpolate_u_p_(buf.get("up") if "up" in buf else vectorfield(setup), u, setup)
            return _interior(setup, buf["up"], True)
        if fieldname == "velocitynorm":
            buf["ke"] = kinetic_energy_(buf.get("ke") if "ke" in buf else scalarfield(setup), u, setup)
            return np.sqrt(2 * _interior(setup, buf["ke"], False))
        if fieldname == "vorticity":
            apply_bc_u_(u, st.get("t", 0.0), setup)
            w = vorticity_(buf.get("w") if "w" in buf else (scalarfield(setup) if D == 2 else vectorfield(setup)), u, setup)
            buf["w"] = w
            wp = interpolate_omega_p_(buf.get("wp") if "wp" in buf else torch.zeros_like(w), w, setup)
            buf["wp"] = wp
            return _interior(setup, wp, D == 3)
        if fieldname == "qcrit":
            buf["q"] = qcrit_(buf.get("q") if "q" in buf else scalarfield(setup), u, setup)
            return _interior(setup, buf["q"], False)
        if fieldname == "temperature":
            return _interior(setup, st["temp"], False)
        raise ValueError("Unknown fieldname")

    field = Observable(compute(state.value))
    state.on(lambda st: setattr(field, "value", compute(st)))
    return field


def observespectrum(state, *, setup, psolver):
    """observespectrum (src/processors.jl:249-277) -> dict(ehat=Observable, kappa=...); the spectrum is
    evaluated by the psolver_spectral plan's FFT passes (insb200_energy_spectrum)."""
    state = state if isinstance(state, Observable) else Observable(state)
    first = energyspectrum(state.value["u"], setup, psolver)
    ehat = Observable(first["ehat"])
    state.on(lambda st: setattr(ehat, "value", energyspectrum(st["u"], setup, psolver)["ehat"]))
    return dict(ehat=ehat, kappa=first["kappa"])


def _write_vtr(filename, coords, fields, t):
    """Minimal VTK XML rectilinear-grid (.vtr) writer, ASCII point data at the pressure points
    (what WriteVTK's vtk_grid(filename, xp...) produces for the reference, processors.jl:163-190)."""
    dims = [len(c) for c in coords] + [1] * (3 - len(coords))
    ext = " ".join(f"0 {n - 1}" for n in dims)

    def arr(name, data, ncomp=1):
        flat = " ".join(repr(float(v)) for v in np.asarray(data, dtype=np.float64).ravel())
        return (f'<DataArray type="Float64" Name="{name}" NumberOfComponents="{ncomp}" format="ascii">{flat}</DataArray>')

    out = ['<?xml version="1.0"?>', '<VTKFile type="RectilinearGrid" version="1.0" byte_order="LittleEndian">',
           f'<RectilinearGrid WholeExtent="{ext}">', '<FieldData>', arr("TimeValue", [t]), '</FieldData>',
           f'<Piece Extent="{ext}">', '<PointData>']
    for name, f in fields.items():
        if isinstance(f, tuple):  # vector: components interleaved per point, x fastest
            comps = [np.asarray(c, dtype=np.float64).ravel(order="F") for c in f]
            out.append(arr(name, np.stack(comps, axis=1), len(comps)))
        else:
            out.append(arr(name, np.asarray(f, dtype=np.float64).ravel(order="F")))
    out += ['</PointData>', '<Coordinates>']
    for i in range(3):
        out.append(arr("xyz"[i], coords[i] if i < len(coords) else [0.0]))
    out += ['</Coordinates>', '</Piece>', '</RectilinearGrid>', '</VTKFile>']
    if not filename.endswith(".vtr"):
        filename += ".vtr"
    with open(filename, "w") as fh:
        fh.write("\n".join(out))
    return filename


def snapshotsaver(state, *, setup, fieldnames=("velocity",), psolver=None):
    """snapshotsaver (src/processors.jl:150-190): returns savesnapshot(filename) writing the observed
    fields at the pressure points to a .vtr file; 2-D vectors get a zero z-component."""
    state = state if isinstance(state, Observable) else Observable(state)
    D = setup.D
    coords = [np.asarray(setup.xp[a][setup.Ip[a][0]:setup.Ip[a][1]], dtype=np.float64) for a in range(D)]
    fields = {fn: observefield(state, setup=setup, fieldname=fn, psolver=psolver) for fn in fieldnames}

    def savesnapshot(filename, pvd=None):
        data = {}
        for fn, f in fields.items():
            g = f.value
            if g.ndim == D:
                data[str(fn)] = g
            else:
                comps = [g[..., c] for c in range(g.shape[-1])]
                if D == 2:
                    comps.append(np.zeros_like(comps[0]))
                data[str(fn)] = tuple(comps)
        path = _write_vtr(filename, coords, data, float(state.value.get("t", 0.0)))
        if pvd is not None:
            pvd.append((float(state.value.get("t", 0.0)), path))
        return path

    return savesnapshot


def save_vtk(state, *, setup, filename="output/solution", **kwargs):
    """save_vtk (src/processors.jl:197-202)."""
    d = os.path.dirname(filename)
    if d:
        os.makedirs(d, exist_ok=True)
    return snapshotsaver(state, setup=setup, **kwargs)(filename)


def vtk_writer(setup, nupdate=1, dir="output", filename="solution", **kwargs):
    """vtk_writer (src/processors.jl:210-230): a .vtr file every nupdate steps + a ParaView .pvd
    collection written by finalize."""

    def initialize(outer):
        os.makedirs(dir, exist_ok=True)
        pvd = []
        inner = Observable(outer.value)
        save = snapshotsaver(inner, setup=setup, **kwargs)

        def cb(st):
            if st["n"] % nupdate:
                return
            inner.value = st
            save(os.path.join(dir, f"{filename}_t={str(float(st['t'])).replace('.', 'p')}"), pvd)

        outer.on(cb)
        cb(outer.value)
        return pvd

    def finalize(pvd, state):
        lines = ['<?xml version="1.0"?>', '<VTKFile type="Collection" version="1.0">', '<Collection>']
        lines += [f'<DataSet timestep="{t!r}" part="0" file="{os.path.basename(p)}"/>' for t, p in pvd]
        lines += ['</Collection>', '</VTKFile>']
        path = os.path.join(dir, filename + ".pvd")
        with open(path, "w") as fh:
            fh.write("\n".join(lines))
        return path

    return processor(initialize, finalize, nupdate=nupdate)


def solve_unsteady(*, setup, tlims, start, force=navierstokes_, docopy=True, method=None, psolver=None,
                   dt=None, dt_min=None, cfl=None, n_adapt_dt=1, processors=None, params, native=True, graph=True):
    """solve_unsteady (src/solver.jl:77-149).  native=True runs the loop inside the library
    (insb200_solve_unsteady); native=False steps from Python, one library call per step."""
    T = setup.T
    method = method if method is not None else LMWray3(T)
    psolver = psolver if psolver is not None else default_psolver(setup)
    cfl = T(0.9) if cfl is None else cfl
    tstart, tend = tlims
    state = {k: (v.clone() if docopy else v) for k, v in start.items()}
    stepper = create_stepper(method, setup=setup, psolver=psolver, state=state, t=tstart)
    obs = Observable(get_state(stepper))
    processors = processors or {}
    initialized = {k: p.initialize(obs) for k, p in processors.items()}
    planes = _bc_planes(setup)
    # (a Dirichlet function not declared time independent keeps being probed step by step: Python loop)
    if native and force in (navierstokes_, boussinesq_) and not planes.time_dependent and not getattr(planes, "probing", False):
        import math

        every = 0
        for p in processors.values():
            every = math.gcd(every, max(int(getattr(p, "nupdate", 1)), 1))
        stepper = _solve_native(setup, method, force, stepper, tend, dt, dt_min, cfl, n_adapt_dt, params, obs, every,
                                graph)
    elif dt is None:
        while stepper["t"] < tend:
            if stepper["n"] % n_adapt_dt == 0:
                dt = cfl * propose_timestep(force, stepper["state"], setup, params, plan=psolver.plan)
                if dt_min is not None:
                    dt = max(dt, dt_min)
                if dt < 1e-10 or dt > 100 or np.isnan(dt):
                    import warnings

                    warnings.warn(f"Proposed time step {dt} is out of bounds. Stopping simulation.")
                    break
            dt_step = min(dt, tend - stepper["t"])
            stepper = timestep_(method, force, stepper, dt_step, params=params)
            obs.value = get_state(stepper)
    else:
        nstep = int(round((tend - tstart) / dt))
        dt = (tend - tstart) / nstep
        for _ in range(nstep):
            stepper = timestep_(method, force, stepper, dt, params=params)
            obs.value = get_state(stepper)
    outputs = {k: processors[k].finalize(initialized[k], obs) for k in processors}
    out = dict(stepper["state"])
    out["t"] = stepper["t"]
    return out, outputs


# ------------------------------------------------------------------------------------
# SciML right-hand side (src/sciml.jl)
# ------------------------------------------------------------------------------------
def right_hand_side_(dudt, u, params_ref, t):
    setup, psolver, viscosity = params_ref
    if _bc_planes(setup).funcs:
        raise NotImplementedError("right_hand_side! with function-valued DirichletBC: compose apply_bc_u_(u, t), "
                                  "navierstokes_, apply_bc_u_(dudt, t, dudt=True) and project_ on the host")
    pr = _params(setup, viscosity=viscosity)
    check(lib().insb200_right_hand_side(psolver.plan.h, _ptr(_chk(setup, dudt, True)), _ptr(_chk(setup, u, True)),
                                        float(t), C.byref(pr), _stream()))
    return None


def create_right_hand_side(setup, psolver):
    def right_hand_side(u, params, t):
        dudt = torch.empty_like(u)
        right_hand_side_(dudt, u, (setup, psolver, params["viscosity"]), t)
        return dudt

    return right_hand_side


# ------------------------------------------------------------------------------------
# initial conditions (src/initializers.jl:45-87): sample on the host, project on the device
# ------------------------------------------------------------------------------------
def velocityfield(setup, ufunc, t=0.0, *, psolver=None, doproject=True):
    D = setup.D
    u = np.zeros(setup.N_local + (D,), dtype=setup.T, order="F")
    for a in range(D):
        xs, sl = [], []
        for b in range(D):
            lo, hi = setup.local_range(setup.Iu[a][b], b)
            off = setup.zoff if (b == 2 and setup.world > 1) else 0
            shape = [1] * D
            shape[b] = hi - lo
            xs.append(setup.xu[a][b][lo + off:hi + off].reshape(shape))
            sl.append(slice(lo, hi))
        u[tuple(sl) + (a,)] = np.asarray(ufunc(a, *xs), dtype=setup.T)
    ud = to_device(setup, u)
    apply_bc_u_(ud, t, setup)
    if doproject:
        psolver = psolver if psolver is not None else default_psolver(setup)
        ud = project(ud, setup, psolver=psolver)
        apply_bc_u_(ud, t, setup)
    return ud


def orlandi_profile(k, kpeak=10):
    """orlandi_profile (src/initializers.jl:93): E(k) ~ k^4 exp(-2 (k / kpeak)^2)."""
    return k**4 * np.exp(-2 * (k / kpeak) ** 2)


def random_field(setup, t=0.0, *, profile=orlandi_profile, totalenergy=1.0, rng=None, psolver=None, **kwargs):
    """random_field (src/initializers.jl:109-183), synthesised on the device with the FFT passes of the
    setup's psolver_spectral (insb200_random_field): divergence free on the staggered grid to round-off, shell
    energies equal to ``totalenergy * profile(kappa) / sum(profile)``.  ``rng``: a NumPy Generator draws the white
    noise on the host like the reference's ``randn!(rng, v)`` per component (reproducible against the oracle);
    ``None`` or an integer seed generates it on the device (Philox4x32-10), so that fields too large for the host
    never exist there.  The profile function is evaluated here (it cannot cross the ABI)."""
    if any(not setup.is_periodic(d) for d in range(setup.D)):
        raise ValueError("Random field requires periodic boundary conditions.")
    T, D, Np = setup.T, setup.D, setup.Np
    psolver = psolver if psolver is not None else psolver_spectral(setup)
    kdiag = int(np.floor(np.sqrt(sum((n // 2) ** 2 for n in Np))))
    while (kdiag + 1) ** 2 <= sum((n // 2) ** 2 for n in Np):
        kdiag += 1
    prof = [T(profile(T(k), **kwargs)) for k in range(kdiag + 1)]
    totalprofile = sum(prof)
    e0 = (C.c_double * (kdiag + 1))(*[float(T(totalenergy) * pk / totalprofile) for pk in prof])
    u = vectorfield(setup)
    if isinstance(rng, np.random.Generator):
        host = np.zeros(setup.N + (D,), dtype=T, order="F")
        for a in range(D):
            host[tuple(slice(lo, hi) for lo, hi in setup.Ip) + (a,)] = rng.standard_normal(Np).astype(T)
        u = to_device(setup, host)
        fill, seed = 0, 0
    else:
        fill, seed = 1, int(rng) if rng is not None else int(np.random.SeedSequence().entropy % (1 << 63))
    check(lib().insb200_random_field(psolver.plan.h, _ptr(u), e0, kdiag + 1, fill, C.c_ulonglong(seed), float(t),
                                     _stream()))
    return u


def temperaturefield(setup, tempfunc, t=0.0):
    D = setup.D
    temp = np.zeros(setup.N_local, dtype=setup.T, order="F")
    xs, sl = [], []
    for b in range(D):
        lo, hi = setup.local_range(setup.Ip[b], b)
        off = setup.zoff if (b == 2 and setup.world > 1) else 0
        shape = [1] * D
        shape[b] = hi - lo
        xs.append(setup.xp[b][lo + off:hi + off].reshape(shape))
        sl.append(slice(lo, hi))
    temp[tuple(sl)] = np.asarray(tempfunc(*xs), dtype=setup.T)
    td = to_device(setup, temp)
    return apply_bc_temp_(td, t, setup)
