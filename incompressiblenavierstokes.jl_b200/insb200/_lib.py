"""ctypes binding of libinsb200.so (include/insb200.h).  No compute happens in Python:
every function below forwards device pointers to the C ABI, and loading fails loudly when
the CUDA library has not been built -- there is no CPU fallback."""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libinsb200.so")

F32, F64 = 0, 1
BC_PERIODIC, BC_DIRICHLET, BC_SYMMETRIC, BC_PRESSURE = 0, 1, 2, 3
PSOLVER_NONE, PSOLVER_SPECTRAL, PSOLVER_TRANSFORM, PSOLVER_TRIDIAGONAL, PSOLVER_DEFAULT, PSOLVER_CG = -1, 0, 1, 2, 3, 4


class SetupStruct(C.Structure):
    _fields_ = [
        ("D", C.c_int),
        ("N", C.c_int * 3),
        ("dtype", C.c_int),
        ("bc_u", (C.c_int * 2) * 3),
        ("has_temp", C.c_int),
        ("bc_temp", (C.c_int * 2) * 3),
        ("Delta", C.c_void_p * 3),
        ("DeltaU", C.c_void_p * 3),
        ("A1", C.c_void_p * 3),
        ("A2", C.c_void_p * 3),
        ("xlims", (C.c_double * 2) * 3),
        ("bc_u_value", ((C.c_double * 3) * 2) * 3),
        ("bc_temp_value", (C.c_double * 2) * 3),
        ("psolver", C.c_int),
        ("tridiagonal_dir", C.c_int),
        ("eig", C.c_void_p * 3),
        ("device", C.c_int),
        ("rank", C.c_int),
        ("world", C.c_int),
        ("nccl_id", C.c_void_p),
        ("cg_abstol", C.c_double),
        ("cg_reltol", C.c_double),
        ("cg_maxiter", C.c_longlong),
    ]


class ParamsStruct(C.Structure):
    _fields_ = [
        ("viscosity", C.c_double),
        ("conductivity", C.c_double),
        ("gravity", C.c_double),
        ("gdir", C.c_int),
        ("dodissipation", C.c_int),
        ("boussinesq", C.c_int),
        ("bodyforce", C.c_double * 3),
        ("eddy_model", C.c_int),
        ("eddy_C", C.c_double),
    ]


# every symbol include/insb200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_PROTOTYPES = {
    "insb200_last_error": (C.c_char_p, []),
    "insb200_version": (C.c_int, []),
    "insb200_plan_create": (C.c_int, [C.POINTER(SetupStruct), C.POINTER(_P)]),
    "insb200_plan_destroy": (C.c_int, [_P]),
    "insb200_comm_unique_id": (C.c_int, [_P]),
    "insb200_plan_local_n": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "insb200_plan_ipc_export": (C.c_int, [_P, _P]),
    "insb200_plan_ipc_import": (C.c_int, [_P, _P]),
    "insb200_plan_ipc_close": (C.c_int, [_P]),
    "insb200_plan_psolver": (C.c_int, [_P]),
    "insb200_plan_workspace_bytes": (C.c_longlong, [_P]),
    "insb200_apply_bc_u": (C.c_int, [_P, _P, C.c_double, C.c_int, _P]),
    "insb200_apply_bc_p": (C.c_int, [_P, _P, _P]),
    "insb200_apply_bc_temp": (C.c_int, [_P, _P, C.c_double, _P]),
    "insb200_set_bc_u_plane": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P]),
    "insb200_rhs": (C.c_int, [_P, _P, _P, _P, _P, C.POINTER(ParamsStruct), _P]),
    "insb200_convection": (C.c_int, [_P, _P, _P, _P]),
    "insb200_diffusion": (C.c_int, [_P, _P, _P, C.c_double, _P]),
    "insb200_divergence": (C.c_int, [_P, _P, _P, _P]),
    "insb200_scalewithvolume": (C.c_int, [_P, _P, _P]),
    "insb200_pressuregradient": (C.c_int, [_P, _P, _P, _P]),
    "insb200_applypressure": (C.c_int, [_P, _P, _P, _P]),
    "insb200_psolve": (C.c_int, [_P, _P, _P]),
    "insb200_project": (C.c_int, [_P, _P, _P, _P]),
    "insb200_stage_combine": (C.c_int, [_P, _P, _P, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_double), C.c_int,
                                        C.c_int, _P]),
    "insb200_stage_combine_project": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.POINTER(C.c_void_p),
                                                C.POINTER(C.c_void_p), C.POINTER(C.c_double), C.c_int, _P, C.c_double,
                                                _P]),
    "insb200_axpy": (C.c_int, [_P, _P, C.c_double, _P, C.c_int, C.c_int, _P]),
    "insb200_step_rk": (
        C.c_int,
        [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_int, _P, _P, C.c_double,
         C.c_double, C.POINTER(ParamsStruct), _P],
    ),
    "insb200_step_lmwray3": (C.c_int, [_P, _P, _P, C.c_double, C.c_double, C.POINTER(ParamsStruct), _P]),
    "insb200_step_rk_host": (
        C.c_int,
        [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_int, _P, _P, C.c_double,
         C.c_double, C.c_int, C.POINTER(ParamsStruct), _P],
    ),
    "insb200_step_rk_host_ensemble": (
        C.c_int,
        [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_int, C.POINTER(C.c_void_p),
         C.POINTER(C.c_void_p), C.c_int, C.c_double, C.c_double, C.c_int, C.POINTER(ParamsStruct), _P],
    ),
    "insb200_psolver_iterations": (C.c_int, [_P]),
    "insb200_eddy_viscosity_closure": (C.c_int, [_P, C.c_int, C.c_double, _P, _P, _P]),
    "insb200_vorticity": (C.c_int, [_P, _P, _P, _P]),
    "insb200_interpolate_u_p": (C.c_int, [_P, _P, _P, _P]),
    "insb200_interpolate_omega_p": (C.c_int, [_P, _P, _P, _P]),
    "insb200_qcrit": (C.c_int, [_P, _P, _P, _P]),
    "insb200_kinetic_energy": (C.c_int, [_P, _P, _P, C.c_int, _P]),
    "insb200_total_kinetic_energy": (C.c_int, [_P, _P, C.c_int, C.POINTER(C.c_double), _P]),
    "insb200_energy_spectrum": (C.c_int, [_P, _P, C.c_int, C.POINTER(C.c_double), _P]),
    "insb200_spectral_moments": (C.c_int, [_P, _P, C.POINTER(C.c_double), _P]),
    "insb200_random_field": (C.c_int, [_P, _P, C.POINTER(C.c_double), C.c_int, C.c_int, C.c_ulonglong, C.c_double, _P]),
    "insb200_solve_unsteady": (
        C.c_int,
        [_P, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_int, _P, _P, C.c_double, C.c_double,
         C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.POINTER(ParamsStruct), C.c_int, _P, _P, C.c_int,
         C.POINTER(C.c_double), C.POINTER(C.c_longlong), _P],
    ),
    "insb200_right_hand_side": (C.c_int, [_P, _P, _P, C.c_double, C.POINTER(ParamsStruct), _P]),
    "insb200_cfl": (C.c_int, [_P, _P, C.POINTER(ParamsStruct), C.POINTER(C.c_double), _P]),
    "insb200_maxabs": (C.c_int, [_P, _P, C.POINTER(C.c_double), _P]),
    "insb200_launch_count": (C.c_longlong, [_P]),
    "insb200_debug_fused_copy": (C.c_int, [_P, C.c_int, _P, C.c_longlong, C.POINTER(C.c_int)]),
    "insb200_profile_enable": (C.c_int, [_P, C.c_int]),
    "insb200_profile_read": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "insb200_profile_read_ex": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
}

EXPORTED_SYMBOLS = tuple(_PROTOTYPES)

_lib = None


class Insb200Error(RuntimeError):
    pass


def lib():
    """Load libinsb200.so (once).  Raises if it was not built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Insb200Error(
                f"{LIB_PATH} not found: build it with `make -C incompressiblenavierstokes.jl_b200/csrc` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback."
            )
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOTYPES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(status):
    if status != 0:
        msg = lib().insb200_last_error()
        raise Insb200Error(msg.decode() if msg else f"libinsb200 status {status}")
