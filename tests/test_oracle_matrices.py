"""test/matrices.jl:1-52 re-run on the oracle: the slicing restatement of the kernels equals
independently assembled sparse matrices (oracle/matrices.py), on the reference's own test grid
(tanh / cosine stretched, periodic x Dirichlet-Pressure x symmetric)."""
import numpy as np
import pytest

import oracle as O
from oracle import bc as obc, matrices as om, operators as oops


def _setup(T=np.float64):
    x = (O.tanh_grid(0.0, 5.0, 11), O.cosine_grid(0.0, 1.0, 7), O.tanh_grid(0.0, 0.8, 5))
    bcs = ((O.PeriodicBC(), O.PeriodicBC()), (O.DirichletBC(), O.PressureBC()), (O.SymmetricBC(), O.SymmetricBC()))
    return O.Setup(tuple(np.asarray(v).astype(T) for v in x), {"u": bcs})


@pytest.fixture
def fields():
    so = _setup()
    rng = np.random.default_rng(0)
    u = np.asfortranarray(rng.standard_normal(so.N + (so.D,)))
    p = np.asfortranarray(rng.standard_normal(so.N))
    return so, u, p


def test_divergence_matrix(fields):
    so, u, p = fields
    ub = obc.apply_bc_u(u.copy(order="F"), 0.0, so)
    d1 = (om.divergence_mat(so) @ ub.ravel(order="F")).reshape(so.N, order="F")
    d2 = oops.divergence_(O.scalarfield(so), ub, so)
    assert np.allclose(d1, d2, rtol=1e-13, atol=1e-13 * np.abs(d2).max())


def test_pressure_gradient_matrix(fields):
    so, u, p = fields
    pb = obc.apply_bc_p(p.copy(order="F"), 0.0, so)
    g1 = (om.pressuregradient_mat(so) @ pb.ravel(order="F")).reshape(so.N + (so.D,), order="F")
    g2 = oops.pressuregradient_(O.vectorfield(so), pb, so)
    assert np.allclose(g1, g2, rtol=1e-13, atol=1e-13 * np.abs(g2).max())


def test_diffusion_matrix(fields):
    so, u, p = fields
    ub = obc.apply_bc_u(u.copy(order="F"), 0.0, so)
    d1 = (om.diffusion_mat(so) @ ub.ravel(order="F")).reshape(so.N + (so.D,), order="F")
    d2 = oops.diffusion_(O.vectorfield(so), ub, so, 1.0)
    assert np.allclose(d1, d2, rtol=1e-11, atol=1e-12 * np.abs(d2).max())


@pytest.mark.parametrize("bcs", [
    ((O.PeriodicBC(), O.PeriodicBC()), (O.DirichletBC(), O.PressureBC()), (O.SymmetricBC(), O.SymmetricBC())),
    ((O.DirichletBC(), O.DirichletBC()),) * 3,
    ((O.PeriodicBC(), O.PeriodicBC()), (O.DirichletBC(), O.DirichletBC()), (O.PeriodicBC(), O.PeriodicBC())),
], ids=["mixed", "cavity", "channel"])
def test_laplacian_is_volume_times_divergence_of_gradient(bcs):
    """laplacian_mat = P' Ω M Bu G Bp P (matrices.jl:484-493) against laplacian! (operators.jl:320-359),
    the operator psolver_cg iterates on."""
    x = (O.tanh_grid(0.0, 5.0, 11), O.cosine_grid(0.0, 1.0, 7), O.tanh_grid(0.0, 0.8, 5))
    so = O.Setup(tuple(np.asarray(v) for v in x), {"u": bcs})
    rng = np.random.default_rng(1)
    p = O.scalarfield(so)
    p[so.Ip] = rng.standard_normal(so.Np)          # P: pad the DOFs
    obc.apply_bc_p(p, 0.0, so)                     # Bp
    g = (om.pressuregradient_mat(so) @ p.ravel(order="F")).reshape(so.N + (so.D,), order="F")
    g = np.asfortranarray(g)
    # Bu: the homogeneous velocity boundary conditions (no-slip walls keep the wall-normal gradient out)
    hom = O.Setup(tuple(np.asarray(v) for v in x), {"u": tuple(
        tuple(O.DirichletBC() if isinstance(b, O.DirichletBC) else b for b in pair) for pair in bcs)})
    obc.apply_bc_u(g, 0.0, hom)
    lap1 = (om.volume_mat(so) @ (om.divergence_mat(so) @ g.ravel(order="F"))).reshape(so.N, order="F")
    lap2 = oops.laplacian_(O.scalarfield(so), p, so)
    assert np.allclose(lap1[so.Ip], lap2[so.Ip], rtol=1e-11, atol=1e-12 * np.abs(lap2).max())
