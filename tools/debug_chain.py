"""Debugging aid: chained native-loop steps case by case with progress prints (run under `timeout`)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "incompressiblenavierstokes.jl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch

import insb200 as I
from gpu_helpers import make_setups, smooth_vec, uniform

T = np.float32 if (len(sys.argv) > 1 and sys.argv[1] == "f32") else np.float64
mname = sys.argv[2] if len(sys.argv) > 2 else "RK44"
so, sg = make_setups((uniform(70), uniform(24), uniform(20)), ("P", "P", "P"), T=T)
ps = I.psolver_spectral(sg)
u0 = I.to_device(sg, smooth_vec(so, 33))
I.project_(u0, sg, psolver=ps, p=I.scalarfield(sg))
I.apply_bc_u_(u0, 0.0, sg)
m = I.LMWray3(T) if mname == "LMWray3" else I.RKMethods.RK44(T)
base = dict(setup=sg, start={"u": u0}, method=m, psolver=ps, params=dict(viscosity=1e-2))
for name, kw in (("fixed", dict(tlims=(0.0, 0.014), dt=0.002)), ("fixedT", dict(tlims=(0.0, 0.014), dt=T(0.002))),
                 ("adaptive", dict(tlims=(0.0, 0.02), cfl=0.4, n_adapt_dt=3))):
    for tag, opts in (("python", dict(native=False)), ("native", dict(native=True, graph=False)),
                      ("graph", dict(native=True, graph=True))):
        print("start", name, tag, flush=True)
        out, _ = I.solve_unsteady(**base, **kw, **opts)
        torch.cuda.synchronize()
        print("  done", name, tag, float(out["t"]), float(out["u"].abs().max()), flush=True)
