"""Known-answer pins of oracle/eddyviscosity.py (the reference ships no test of src/eddyviscosity.jl):
closed forms of the four models for a velocity field with a constant gradient tensor (the staggered
differences are exact for linear fields), and the closure term of a parabolic shear flow, which the
discrete operators reproduce exactly on a uniform grid (central differences of quadratics)."""
import numpy as np
import pytest

import oracle as O
from oracle import eddyviscosity as oev

from gpu_helpers import _bc_pair, uniform


def linear_field(so, A):
    """u_a = sum_b A[a, b] x_b sampled at the velocity points of component a."""
    u = O.vectorfield(so)
    for a in range(3):
        acc = 0
        for b in range(3):
            shape = [1, 1, 1]
            shape[b] = -1
            acc = acc + A[a, b] * np.asarray(so.xu[a][b], dtype=np.float64).reshape(shape)
        u[..., a] = acc
    return u


def closed_form(model, C, A, delta, d2):
    """nu_t of the model for the gradient tensor A (A[a, b] = du_a/dx_b), filter width delta = (dx dy dz)^(1/3),
    d2 = (dx^2, dy^2, dz^2).  Written with matrix algebra, independently of the oracle's component code."""
    S = 0.5 * (A + A.T)
    if model == "smagorinsky":  # (C delta)^2 sqrt(2 S:S)
        return (C * delta) ** 2 * np.sqrt(2 * np.sum(S * S))
    if model == "wale":  # Nicoud & Ducros 1999
        G2 = A @ A
        Sd = 0.5 * (G2 + G2.T) - np.trace(G2) / 3 * np.eye(3)
        sdsd, ss = np.sum(Sd * Sd), np.sum(S * S)
        return (C * delta) ** 2 * sdsd**1.5 / (ss**2.5 + sdsd**1.25 + np.finfo(float).eps)
    if model == "qr":  # Verstappen 2011: |r| / q with q = tr(S^2)/2, r = -det S = -tr(S^3)/3
        q = np.trace(S @ S) / 2
        r = -np.linalg.det(S)
        return (C * delta) ** 2 * abs(r) / (q + np.finfo(float).eps)
    if model == "vreman":  # Vreman 2004: alpha[i, j] = du_j/dx_i, beta = alpha^T diag(d2) alpha
        alpha = A.T
        beta = alpha.T @ np.diag(d2) @ alpha
        bb = (beta[0, 0] * beta[1, 1] - beta[0, 1] ** 2 + beta[0, 0] * beta[2, 2] - beta[0, 2] ** 2
              + beta[1, 1] * beta[2, 2] - beta[1, 2] ** 2)
        return C**2 * np.sqrt(bb / np.sum(alpha * alpha))
    raise ValueError(model)


@pytest.mark.parametrize("model,C", [("smagorinsky", 0.17), ("wale", 0.569), ("vreman", 0.28), ("qr", 0.092)])
def test_viscosity_of_a_constant_gradient_matches_the_closed_form(model, C):
    rng = np.random.default_rng(7)
    A = rng.standard_normal((3, 3))
    A -= np.trace(A) / 3 * np.eye(3)  # divergence-free, so that tr S = 0 and det S = tr(S^3)/3
    x = (uniform(8, 1.0), uniform(10, 2.5), uniform(6, 0.9))
    so = O.Setup(tuple(np.asarray(v) for v in x), {"u": tuple(_bc_pair("D", O) for _ in range(3))})
    u = linear_field(so, A)
    G = oev.tensorfield(so)
    oev.gradient_tensor_(G, u, so)
    # interior of the interior: the collocation reads one cell to the left, which is a ghost entry at the boundary
    inner = (slice(2, -2),) * 3
    for a in range(3):
        for b in range(3):
            np.testing.assert_allclose(G["xyz"[a] + "xyz"[b]][inner], A[a, b], rtol=0, atol=1e-12)
    visc = O.scalarfield(so)
    oev.eddy_viscosity_(model, C, visc, G, so)
    h = [1.0 / 8, 2.5 / 10, 0.9 / 6]
    want = closed_form(model, C, A, np.cbrt(h[0] * h[1] * h[2]), np.array(h) ** 2)
    assert want > 1e-6
    np.testing.assert_allclose(visc[inner], want, rtol=2e-11)


def test_smagorinsky_closure_of_a_parabolic_shear_flow_is_exact():
    """u = (y^2/2, 0, 0) on a uniform grid, y > 0: nu_t = (C delta)^2 y, sigma_xy = -(C delta)^2 y^2 at the
    edges, closure term f_x = -d(sigma_xy)/dy = 2 (C delta)^2 y at the velocity points, f_y = f_z = 0."""
    C = 0.2
    x = (uniform(6, 1.5), 0.5 + uniform(12, 3.0), uniform(5, 1.0))
    so = O.Setup(tuple(np.asarray(v) for v in x), {"u": (_bc_pair("P", O), _bc_pair("D", O), _bc_pair("P", O))})
    u = O.vectorfield(so)
    y_u = np.asarray(so.xu[0][1], dtype=np.float64).reshape(1, -1, 1)
    u[..., 0] = 0.5 * y_u**2 + 0 * u[..., 0]
    f = oev.eddy_viscosity_closure_("smagorinsky", C, O.vectorfield(so), u, so)
    delta = np.cbrt((1.5 / 6) * (3.0 / 12) * (1.0 / 5))
    inner = (slice(1, -1), slice(3, -3), slice(1, -1))  # away from the walls, where the ghost rules enter
    want = 2 * (C * delta) ** 2 * y_u[:, 3:-3, :]
    np.testing.assert_allclose(f[inner + (0,)], np.broadcast_to(want, f[inner + (0,)].shape), rtol=1e-11)
    assert np.abs(f[inner + (1,)]).max() < 1e-13 and np.abs(f[inner + (2,)]).max() < 1e-13
