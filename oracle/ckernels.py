"""Oracle (test infrastructure): ctypes binding of oracle/c/ins_ref_kernels.c, the C/OpenMP
cell-by-cell restatement of the reference's threaded CPU kernels (src/operators.jl).

Used to cross-check the NumPy restatement (tests/test_oracle_ckernels.py) and as the
multithreaded CPU baseline of bench.py.  ``oracle.operators.BACKEND = "c"`` routes the stencil
operators here; the default backend stays NumPy."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libinsref.so")
_lib = None


def build():
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "c")], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.insref_max_threads.restype = C.c_int
    return _lib


def available():
    try:
        lib()
        return True
    except Exception:
        return False


def set_threads(n):
    lib().insref_set_threads(int(n))


def max_threads():
    return int(lib().insref_max_threads())


def _sfx(T):
    return "f32" if np.dtype(T) == np.float32 else "f64"


def _tables(setup, *names):
    """(pointer-array, keep-alive) per table name: REAL *tab[3]"""
    out, keep = [], []
    for nm in names:
        arrs = [np.ascontiguousarray(getattr(setup, nm)[a], dtype=setup.T) for a in range(setup.D)]
        keep.append(arrs)
        ptrs = (C.c_void_p * 3)(*[a.ctypes.data for a in arrs], *([None] * (3 - setup.D)))
        out.append(ptrs)
    return out, keep


def _N(setup):
    return (C.c_int * 3)(*setup.N, *([1] * (3 - setup.D)))


def _real(setup, v):
    return C.c_float(v) if setup.T is np.float32 else C.c_double(v)


def _chk(setup, a):
    assert a.dtype == setup.T and a.flags.f_contiguous, "oracle arrays are Fortran-ordered in the setup dtype"
    return C.c_void_p(a.ctypes.data)


def convdiff_add(f, u, setup, visc, conv=True, diff=True):
    (d, du, a1, a2), keep = _tables(setup, "Delta", "DeltaU", "A1", "A2")
    getattr(lib(), f"insref_convdiff_add_{_sfx(setup.T)}")(
        _chk(setup, f), _chk(setup, u), _N(setup), C.c_int(setup.D), d, du, a1, a2, _real(setup, setup.T(visc)),
        C.c_int(int(conv)), C.c_int(int(diff)))
    return f


def divergence(div, u, setup):
    (d, du), keep = _tables(setup, "Delta", "DeltaU")
    getattr(lib(), f"insref_divergence_{_sfx(setup.T)}")(_chk(setup, div), _chk(setup, u), _N(setup), C.c_int(setup.D), d, du)
    return div


def scalewithvolume(p, setup):
    (d,), keep = _tables(setup, "Delta")
    getattr(lib(), f"insref_scalewithvolume_{_sfx(setup.T)}")(_chk(setup, p), _N(setup), C.c_int(setup.D), d)
    return p


def applypressure(u, p, setup):
    (du,), keep = _tables(setup, "DeltaU")
    getattr(lib(), f"insref_applypressure_{_sfx(setup.T)}")(_chk(setup, u), _chk(setup, p), _N(setup), C.c_int(setup.D), du)
    return u
