#!/usr/bin/env python
"""A/B of the Boussinesq temperature pass on cfg 4 (RayleighBenard3D 512x512x256, RK44), one process: the switches are
read per launch.  tile kernel with / without the operand prefetch (INSB200_BOUSS_PREFETCH=0), Float64 also the one-thread-per-cell
pass (INSB200_BOUSS_TILE=0).

    python tools/bench_bouss.py [--scale 1.0]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "incompressiblenavierstokes.jl_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench_configs import run  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    a = ap.parse_args()
    import insb200 as I

    P, D, S = I.PeriodicBC, I.DirichletBC, I.SymmetricBC
    sz = lambda n: max(8, int(round(n * a.scale / 2)) * 2)
    variants = {
        np.float32: [("tile + prefetch + shared metric tables (default)", {}),
                     ("tile + prefetch", {"INSB200_BOUSS_SMEMTAB": "0"}), ("tile, no prefetch", {"INSB200_BOUSS_PREFETCH": "0"})],
        np.float64: [("tile + prefetch + shared metric tables (default)", {}),
                     ("tile + prefetch", {"INSB200_BOUSS_SMEMTAB": "0"}), ("tile, no prefetch", {"INSB200_BOUSS_PREFETCH": "0"}),
                     ("direct", {"INSB200_BOUSS_TILE": "0"})],
    }
    for T in (np.float32, np.float64):
        n = (sz(512), sz(512), sz(256))
        x = (np.linspace(0, np.pi, n[0] + 1).astype(T), np.linspace(0, 1, n[1] + 1).astype(T),
             np.linspace(0, 1, n[2] + 1).astype(T))
        sg = I.Setup(x, {"u": ((P(), P()), (D(), D()), (D(), D())), "temp": ((P(), P()), (S(), S()), (D(1.0), D(0.0)))})
        ps = I.psolver_transform(sg)
        u = I.vectorfield(sg)
        temp = I.temperaturefield(sg, lambda x, y, z: 0.5 + np.sin(20 * x) * np.sin(np.pi * 20 * y) / 100 + 0 * z)
        params = dict(viscosity=2.5e-4, conductivity=2.5e-4, gravity=1.0, gdir=2, dodissipation=True)
        for tag, env in variants[T]:
            for k in ("INSB200_BOUSS_TILE", "INSB200_BOUSS_PREFETCH", "INSB200_BOUSS_SMEMTAB"):
                os.environ.pop(k, None)
            os.environ.update(env)
            run(f"cfg4 RK44 {np.dtype(T).name} {tag}", sg, ps, {"u": u.clone(), "temp": temp.clone()}, I.RKMethods.RK44(T),
                I.boussinesq_, params, T(1e-3), 5)
        I.release_plans(sg)
        del sg, ps, u, temp
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
