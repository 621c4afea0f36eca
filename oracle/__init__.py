"""CPU oracle for the insb200 hot path -- TEST INFRASTRUCTURE ONLY.

This package is a NumPy restatement of the per-time-step hot path of
agdestein/IncompressibleNavierStokes.jl v5.0.1 (reference checkout at
/root/reference).  Every function cites the reference file:line it follows and
keeps the reference's evaluation order (no fusion, no re-association), so that
it can stand in for the reference's CPU path, which cannot be executed here
(Julia is not installed in this image and there is no network).

Pinning status
--------------
* Pinned against the reference's OWN known-answer tests, ported to
  ``tests/test_oracle_*.py``: manufactured pressure (test/psolvers.jl:13-36),
  tridiagonal channel solver (test/psolvers.jl:39-83), manual DCT == library
  DCT (test/dct.jl:14-44), D = -G^T (test/operators.jl:54-84), convection
  skew-symmetry 1e-12 (test/operators.jl:103-124), diffusion dissipative
  (:126-150), convectiondiffusion == convection + diffusion (:152-167),
  pulsating plug flow 1e-10 (test/timesteppers.jl:44-91), analytic 2-D
  Taylor-Green decay (examples/TaylorGreenVortex2D.jl:59-60), random_field
  divergence-free (test/spectral.jl:29-69).
* Bit-level parity against a Julia build is UNPINNED: the reference ships no
  golden vectors on disk and cannot be run here.  FFTs come from
  scipy.fft (pocketfft) instead of FFTW.jl (not vendored, version unpinned by
  the reference); results agree with any correct FFT to round-off.

Who may import this package: ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.  The product
(``incompressiblenavierstokes.jl_b200/``) never imports it.
"""

from .grid import (  # noqa: F401
    PeriodicBC,
    DirichletBC,
    SymmetricBC,
    PressureBC,
    Setup,
    cosine_grid,
    stretched_grid,
    tanh_grid,
    scalarfield,
    vectorfield,
)
from . import bc, operators, pressure, timestep, initializers  # noqa: F401
