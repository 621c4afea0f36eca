"""Oracle (test infrastructure): sparse-matrix forms of the linear operators, assembled entry by
entry from the kernel definitions the way /root/reference/src/matrices.jl:385-557 does
(divergence_mat, pressuregradient_mat, volume_mat, diffusion_mat).  They give the slicing
restatement in operators.py an independently structured cross-check -- the reference validates
its own kernels against these matrices in test/matrices.jl:18-52."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def _lin(shape):
    """column-major linear indices (the reference's reshape(1:n, N...))"""
    return np.arange(int(np.prod(shape))).reshape(shape, order="F")


def _interior_index_grids(N):
    return np.meshgrid(*[np.arange(1, n - 1) for n in N], indexing="ij")


def divergence_mat(setup):
    """matrices.jl:385-427: div[I] += u[I, i] / Δ_i[I_i] - u[I - e_i, i] / Δ_i[I_i]."""
    D, N = setup.D, setup.N
    n = int(np.prod(N))
    ip, iu = _lin(N), _lin(N + (D,))
    I = _interior_index_grids(N)
    rows, cols, vals = [], [], []
    for i in range(D):
        d = setup.Delta[i][I[i]]
        left = list(I)
        left[i] = I[i] - 1
        rows += [ip[tuple(I)].ravel(), ip[tuple(I)].ravel()]
        cols += [iu[tuple(I) + (i,)].ravel(), iu[tuple(left) + (i,)].ravel()]
        vals += [(1 / d).ravel(), (-1 / d).ravel()]
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n * D))


def pressuregradient_mat(setup):
    """matrices.jl:430-469: G[I, i] = p[I + e_i] / Δu_i[I_i] - p[I] / Δu_i[I_i]."""
    D, N = setup.D, setup.N
    n = int(np.prod(N))
    ip, iu = _lin(N), _lin(N + (D,))
    I = _interior_index_grids(N)
    rows, cols, vals = [], [], []
    for i in range(D):
        d = setup.DeltaU[i][I[i]]
        right = list(I)
        right[i] = I[i] + 1
        rows += [iu[tuple(I) + (i,)].ravel(), iu[tuple(I) + (i,)].ravel()]
        cols += [ip[tuple(right)].ravel(), ip[tuple(I)].ravel()]
        vals += [(1 / d).ravel(), (-1 / d).ravel()]
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n * D, n))


def volume_mat(setup):
    """matrices.jl:472-478: diagonal of the cell volumes (the whole ghosted array)."""
    D, N = setup.D, setup.N
    w = setup.Delta[0].reshape((-1,) + (1,) * (D - 1))
    for a in range(1, D):
        shape = [1] * D
        shape[a] = N[a]
        w = w * setup.Delta[a].reshape(shape)
    return sp.diags(np.broadcast_to(w, N).ravel(order="F"))


def diffusion_mat(setup):
    """matrices.jl:495-557: F[I, i] += ((u[I+e_j, i] - u[I, i]) / Δb - (u[I, i] - u[I-e_j, i]) / Δa) / Δ_ij
    with the first-derivative entries dropped where the spacing is <= 2 eps."""
    D, N, T = setup.D, setup.N, setup.T
    n = int(np.prod(N)) * D
    il = _lin(N + (D,))
    I = _interior_index_grids(N)
    tol = 2 * np.finfo(T).eps
    rows, cols, vals = [], [], []
    for i in range(D):
        for j in range(D):
            dab = setup.DeltaU[j][I[j]] if i == j else setup.Delta[j][I[j]]
            da = setup.Delta[j][I[j]] if i == j else setup.DeltaU[j][I[j] - 1]
            db = setup.Delta[j][I[j] + 1] if i == j else setup.DeltaU[j][I[j]]
            left, right = list(I), list(I)
            left[j], right[j] = I[j] - 1, I[j] + 1
            aa = np.where(da > tol, 1 / da / dab, 0.0)
            bb = np.where(db > tol, 1 / db / dab, 0.0)
            me = il[tuple(I) + (i,)].ravel()
            rows += [me, me, me]
            cols += [il[tuple(left) + (i,)].ravel(), il[tuple(right) + (i,)].ravel(), me]
            vals += [aa.ravel(), bb.ravel(), (-(aa + bb)).ravel()]
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
