#!/usr/bin/env python
"""Times eddy_viscosity_closure! alone on the ChannelFlow grid (cfg 5, optionally scaled), fused (default: tile
kernels; INSB200_LES_TILE=0: one thread per cell) against the four-pass path (INSB200_LES_UNFUSED=1), per model and dtype; prints one JSON line per case with ms per closure
and the effective bandwidth over the fused path's 10 scalar fields (u 3 R, nu_t 1 W + 1 R, u 3 R again, f 3 RMW = 6).

    python tools/bench_les.py [--scale 0.5] [--models WALE Smagorinsky] [--reps 5]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "incompressiblenavierstokes.jl_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.5)
    ap.add_argument("--models", nargs="+", default=["WALE", "Smagorinsky"])
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    import insb200 as I

    P, D = I.PeriodicBC, I.DirichletBC
    n = tuple(max(8, int(round(v * a.scale / 2)) * 2) for v in (1024, 512, 512))
    for T in (np.float64, np.float32):
        x = (np.linspace(0, 4 * np.pi, n[0] + 1).astype(T), I.tanh_grid(0.0, 2.0, n[1], 1.4).astype(T),
             np.linspace(0, 4 * np.pi / 3, n[2] + 1).astype(T))
        sg = I.Setup(x, {"u": ((P(), P()), (D(), D()), (P(), P()))})
        g = torch.Generator(device="cuda").manual_seed(1)
        u = I.vectorfield(sg)
        u.copy_(torch.randn(u.shape, generator=g, device="cuda", dtype=u.dtype))
        I.apply_bc_u_(u, 0.0, sg)
        f = I.vectorfield(sg)
        cells = float(np.prod(sg.Np))
        for model in a.models:
            ref = None
            for mode in ("four-pass", "fused, one thread per cell", "fused, no output prefetch", "fused"):
                unfused = mode == "four-pass"
                for k in ("INSB200_LES_UNFUSED", "INSB200_LES_TILE", "INSB200_LES_PREFETCH"):
                    os.environ.pop(k, None)
                if unfused:
                    os.environ["INSB200_LES_UNFUSED"] = "1"
                elif mode == "fused, one thread per cell":
                    os.environ["INSB200_LES_TILE"] = "0"
                elif mode == "fused, no output prefetch":
                    os.environ["INSB200_LES_PREFETCH"] = "0"
                f.zero_()
                I.eddy_viscosity_closure_(getattr(I, model)(0.3), f, u, None, sg)
                out = f.clone()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(a.reps):
                    I.eddy_viscosity_closure_(getattr(I, model)(0.3), f, u, None, sg)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / a.reps
                if unfused:
                    ref = out
                    diff = None
                else:
                    diff = float((out - ref).norm() / ref.norm())
                print(json.dumps({"case": f"closure {model} {np.dtype(T).name} {mode}",
                                  "N": list(sg.N), "ms_per_closure": round(ms, 3),
                                  "GBps_over_10_fields": round(10 * cells * np.dtype(T).itemsize / ms / 1e6, 1),
                                  "rel_diff_vs_four_pass": diff}), flush=True)
                del out
        I.release_plans(sg)
        del sg, u, f
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
