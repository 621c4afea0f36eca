#!/usr/bin/env python
"""One step of the secondary workloads at reduced size, for `ncu -k regex:...` captures of the kernels
the headline bench never launches: Boussinesq temperature pass (cfg 4 shape), Thomas sweep and LES
closure (cfg 5 shape), psolver_cg (stretched cavity).

    ncu --metrics <time, dram bytes, throughputs, occupancy, registers> --clock-control none \\
        -k regex:'bouss|les_|cg_|thomas' -c 40 -o gpurun_out/secondary python tools/profile_secondary.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "incompressiblenavierstokes.jl_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import insb200 as I  # noqa: E402

P, D, S = I.PeriodicBC, I.DirichletBC, I.SymmetricBC
T = np.float32
# cfg 4 shape / 2: Boussinesq, psolver_transform
n = (256, 256, 128)
x = (np.linspace(0, np.pi, n[0] + 1).astype(T), np.linspace(0, 1, n[1] + 1).astype(T), np.linspace(0, 1, n[2] + 1).astype(T))
sg = I.Setup(x, {"u": ((P(), P()), (D(), D()), (D(), D())), "temp": ((P(), P()), (S(), S()), (D(1.0), D(0.0)))})
ps = I.psolver_transform(sg)
st = I.create_stepper(I.RKMethods.RK44(T), setup=sg, psolver=ps,
                      state={"u": I.vectorfield(sg), "temp": I.temperaturefield(sg, lambda x, y, z: 0.5 + np.sin(20 * x) * np.sin(np.pi * 20 * y) / 100 + 0 * z)},
                      t=0.0)
I.timestep_(I.RKMethods.RK44(T), I.boussinesq_, st, T(1e-3),
            params=dict(viscosity=2.5e-4, conductivity=2.5e-4, gravity=1.0, gdir=2, dodissipation=True))
del sg, ps, st
# cfg 5 shape / 4: channel, psolver_tridiagonal, WALE closure, Float64
n = (256, 128, 128)
x = (np.linspace(0, 4 * np.pi, n[0] + 1), I.tanh_grid(0.0, 2.0, n[1], 1.4), np.linspace(0, 4 * np.pi / 3, n[2] + 1))
sg = I.Setup(x, {"u": ((P(), P()), (D(), D()), (P(), P()))})
ps = I.default_psolver(sg)
u = I.velocityfield(sg, lambda d, x, y, z: ((1 - (y - 1) ** 2) * (1 + 0.1 * np.sin(3 * z) * np.cos(2 * x)) if d == 0
                                            else 0.05 * np.sin(2 * x) * np.sin(np.pi * y) * np.cos(3 * z)) + 0 * (x + y + z), psolver=ps)
st = I.create_stepper(I.RKMethods.RK44(), setup=sg, psolver=ps, state={"u": u}, t=0.0)
I.timestep_(I.RKMethods.RK44(), I.navierstokes_, st, 1e-4,
            params=dict(viscosity=1 / 180, bodyforce=(1.0, 0.0, 0.0), eddyviscosity=I.WALE(0.569)))
del sg, ps, st, u
# stretched all-wall cavity: psolver_cg
x = (I.cosine_grid(0.0, 1.0, 96),) * 3
sg = I.Setup(x, {"u": ((D(), D()), (D(), D((1.0, 0.0, 0.0))), (D(), D()))})  # lid-driven cavity
ps = I.default_psolver(sg)
u = I.vectorfield(sg)
I.apply_bc_u_(u, 0.0, sg)
I.project_(u, sg, psolver=ps, p=I.scalarfield(sg))
torch.cuda.synchronize()
print("cg iterations", ps.iterations)
