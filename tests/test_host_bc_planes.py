"""Host logic (no GPU): the planes of a function-valued DirichletBC that the Python mirror hands to
insb200_set_bc_u_plane hold the reference's values at the positions the ghost-fill kernel reads
(bc_plane_kernel: element (i1, i2) of the remaining dimensions at i1 + N[d1] * i2)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "incompressiblenavierstokes.jl_b200"))

import oracle as O  # noqa: E402
from oracle import bc as obc  # noqa: E402


@pytest.mark.parametrize("D", [2, 3])
@pytest.mark.parametrize("dudt", [False, True])
def test_function_bc_planes_match_reference_fill(D, dudt):
    import insb200.grid as G

    def f(dim, *a):
        *xs, t = a
        val = (dim + 1) * np.cos(3 * t)
        for g, x in enumerate(xs):
            val = val + (g + 2) * np.sin((g + 1) * x + dim) * (1 + t)
        return val

    x = (np.linspace(0, 1, 6), O.cosine_grid(0, 2, 7), np.linspace(0, 3, 5))[:D]
    kinds_o = [(O.DirichletBC(f), O.DirichletBC(f)), (O.DirichletBC(f), O.PressureBC()), (O.SymmetricBC(), O.DirichletBC(f))][:D]
    kinds_g = [(G.DirichletBC(f), G.DirichletBC(f)), (G.DirichletBC(f), G.PressureBC()), (G.SymmetricBC(), G.DirichletBC(f))][:D]
    so = O.Setup(tuple(np.asarray(v) for v in x), {"u": tuple(kinds_o)})
    sg = G.Setup(tuple(np.asarray(v) for v in x), {"u": tuple(kinds_g)})
    t = 0.37
    planes = sg.evaluate_bc_planes(t, dudt)
    assert len(planes) == D * sum(1 for pair in kinds_g for b in pair if isinstance(b, G.DirichletBC))
    N = so.N
    for (beta, side, a), pl in planes.items():
        # reference fill of this side alone, on a NaN field: only the Dirichlet plane gets values
        u = np.full(N + (D,), np.nan, order="F")
        obc._bc_u_side(so.bc_u[beta][side], u, beta, t, so, bool(side), dudt)
        lo, hi = sg.Iu[a][beta]
        i = hi if side else lo - 1
        others = [g for g in range(D) if g != beta]
        flat = pl.ravel()
        for idx in np.ndindex(*[N[g] for g in others]):
            I = [0] * D
            I[beta] = i
            for g, v in zip(others, idx):
                I[g] = v
            q = idx[0] + (N[others[0]] * idx[1] if D == 3 else 0)
            want = u[tuple(I) + (a,)]
            assert np.isfinite(want)
            assert flat[q] == pytest.approx(want, rel=1e-12, abs=1e-12 if not dudt else 1e-6)
