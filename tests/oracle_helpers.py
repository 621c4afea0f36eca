"""Shared helpers for the oracle tests (test infrastructure)."""
import numpy as np

import oracle as O
from oracle import bc, operators as ops


def periodic_setup(n, D, T=np.float64, L=1.0, temp=False):
    n = (n,) * D if np.isscalar(n) else tuple(n)
    L = (L,) * D if np.isscalar(L) else tuple(L)
    x = tuple(np.linspace(0, L[a], n[a] + 1).astype(T) for a in range(D))
    b = (O.PeriodicBC(), O.PeriodicBC())
    bcs = dict(u=(b,) * D)
    if temp:
        bcs["temp"] = (b,) * D
    return O.Setup(x, bcs)


def psolver_cg(setup):
    """Test-only stand-in for the reference's psolver_direct (out of scope for
    the product): matrix-free CG on -L, L = laplacian! (operators.jl:320-359)."""
    from scipy.sparse.linalg import LinearOperator, cg
    from oracle.grid import PressureBC

    Ip = setup.Ip
    shape = tuple(s.stop - s.start for s in Ip)
    n = int(np.prod(shape))
    singular = not any(isinstance(b, PressureBC) for pair in setup.bc_u for b in pair)

    def mv(x):
        p = O.scalarfield(setup)
        p[Ip] = x.reshape(shape, order="F")
        bc.apply_bc_p(p, 0.0, setup)
        return -ops.laplacian_(O.scalarfield(setup), p, setup)[Ip].reshape(-1, order="F")

    A = LinearOperator((n, n), matvec=mv, dtype=np.float64)

    def psolve(p):
        f = -p[Ip].reshape(-1, order="F")
        if singular:
            f = f - f.mean()
        x, info = cg(A, f, rtol=1e-15, atol=0, maxiter=20000)
        if singular:
            x = x - x.mean()
        p[Ip] = x.reshape(shape, order="F")
        return p

    return psolve


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300)
