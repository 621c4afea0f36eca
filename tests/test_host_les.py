"""The cell functions of the fused eddy-viscosity kernels (csrc/les_math.cuh) run on the CPU
(tests/host/host_les.cu compiles the same header for the host) against oracle/eddyviscosity.py:
viscosity straight from the velocity, stress divergence straight from velocity + viscosity, with the
ghost-plane rules that replace zero_out_wall! + apply_bc_p! on the stored tensors
(src/eddyviscosity.jl:344-378).  Needs nvcc, no GPU."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle as O
from oracle import bc as obc, eddyviscosity as oev
from oracle.grid import DirichletBC, PeriodicBC, SymmetricBC

from gpu_helpers import _bc_pair, smooth_vec, uniform

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host", "host_les.cu")
HDR = os.path.join(HERE, "..", "incompressiblenavierstokes.jl_b200", "csrc", "les_math.cuh")
LIB = os.path.join(HERE, "_build", "libhostles.so")


@pytest.fixture(scope="module")
def lib():
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets",
                               "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", SRC, "-o", LIB])
    return ctypes.CDLL(LIB)


CASES = [
    ("periodic", (uniform(9), uniform(7, 2.0), uniform(8, 0.5)), ("P", "P", "P")),
    ("channel", (uniform(8, 4.0), O.tanh_grid(0, 2, 10, 1.4), uniform(6, 2.0)), ("P", "D", "P")),
    ("walls", (O.tanh_grid(0, 1, 8, 1.2), O.tanh_grid(0, 1, 7, 1.1), O.cosine_grid(0, 1, 9)), ("D", "D", "D")),
    ("mixed", (uniform(7), O.cosine_grid(0, 1, 8), uniform(6)), ("S", "D", "P")),
    ("sides", (O.tanh_grid(0, 1, 6, 1.3), uniform(5), O.cosine_grid(0, 1, 7)), (("S", "D"), ("D", "S"), "S")),
    ("thin", (uniform(1), uniform(2), uniform(5)), ("P", "S", "D")),
    # periodic directions on stretched grids: the ghost entries of the metric tables are not periodic images
    ("stretched", (O.tanh_grid(0, 1, 6, 1.3), O.cosine_grid(0, 2, 7), O.tanh_grid(0, 1, 5, 1.1)), ("P", "P", "P")),
]
MODELS = [("smagorinsky", 1, 0.17), ("wale", 2, 0.569), ("vreman", 3, 0.28), ("qr", 4, 0.092)]


def tables(so):
    """dl, 1/dl, 1/du, wa, wb per dimension, evaluated in T like build_grid (csrc/api.cu)."""
    T = so.T
    out = []
    for a in range(3):
        dl, du = np.asarray(so.Delta[a], dtype=T), np.asarray(so.DeltaU[a], dtype=T)
        wa, wb = np.full(len(dl), T(0.5)), np.full(len(dl), T(0.5))
        wa[:-1] = dl[1:] / (dl[:-1] + dl[1:])
        wb[:-1] = dl[:-1] / (dl[:-1] + dl[1:])
        out.append([np.ascontiguousarray(v, dtype=T) for v in (dl, T(1) / dl, T(1) / du, wa, wb)])
    return out


def edges(so):
    """BC codes of include/insb200.h per dimension and side."""
    code = {PeriodicBC: 0, DirichletBC: 1, SymmetricBC: 2}
    lo = [code[type(so.bc_u[d][0])] for d in range(3)]
    hi = [code[type(so.bc_u[d][1])] for d in range(3)]
    return lo, hi


def host_closure(lib, so, model, C, u, f, edge_all=False):
    T = so.T
    tabs = tables(so)
    flat = [t for row in tabs for t in row]
    ptrs = (ctypes.c_void_p * 15)(*[t.ctypes.data for t in flat])
    N = (ctypes.c_int * 3)(*so.N)
    lo, hi = edges(so)
    lo_c, hi_c = (ctypes.c_int * 3)(*lo), (ctypes.c_int * 3)(*hi)
    visc = np.zeros(so.N, dtype=T, order="F")
    args = lambda phase: (int(T is np.float64), phase, int(edge_all), N, ptrs, lo_c, hi_c, model, ctypes.c_double(C),
                          ctypes.c_void_p(u.ctypes.data), ctypes.c_void_p(visc.ctypes.data), ctypes.c_void_p(f.ctypes.data))
    assert lib.host_les(*args(1)) == 0
    oev.zero_out_wall_(visc, so)
    obc.apply_bc_p(visc, T(0), so)
    assert lib.host_les(*args(2)) == 0
    return visc


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("mo,mid,C", MODELS)
@pytest.mark.parametrize("name,x,bcs", CASES)
def test_fused_cell_functions_match_oracle(lib, name, x, bcs, mo, mid, C, T):
    xs = tuple(np.asarray(v).astype(T) for v in x)
    so = O.Setup(xs, {"u": tuple(_bc_pair(k, O) for k in bcs)})
    u = smooth_vec(so, 101)
    f0 = smooth_vec(so, 102)
    cache = dict(visc=O.scalarfield(so), G=oev.tensorfield(so))
    want = oev.eddy_viscosity_closure_(mo, C, f0.copy(order="F"), u, so, cache)
    got = f0.copy(order="F")
    visc = host_closure(lib, so, mid, C, np.asfortranarray(u), got)
    # the ghost rules applied to every cell (they reduce to the plain stencil away from the ghost planes)
    got2 = f0.copy(order="F")
    visc2 = host_closure(lib, so, mid, C, np.asfortranarray(u), got2, edge_all=True)
    assert np.array_equal(visc, visc2) and np.array_equal(got, got2)
    tol = 1e-12 if T is np.float64 else 1e-5
    vs = max(np.abs(cache["visc"]).max(), 1e-300)
    assert np.abs(visc - cache["visc"]).max() / vs < tol, (name, mo)
    term_o, term_g = want - f0, got - f0
    scale = max(np.linalg.norm(term_o.ravel()), 1e-300)
    err = np.linalg.norm((term_g - term_o).ravel()) / scale
    assert err < tol, (name, mo, err)
    # ghost cells of f are never touched
    inner = (slice(1, -1),) * 3
    mask = np.ones(f0.shape, bool)
    mask[inner] = False
    assert np.array_equal(got[mask], f0[mask])
