#!/usr/bin/env python
"""Per-pass timing of the spectral Poisson solve on one GPU (CUDA events inside the library).

    python tools/bench_poisson.py --grid 512 512 512 [--dtype f32] [--reps 20]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "incompressiblenavierstokes.jl_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, nargs=3, default=[512, 512, 512])
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    import insb200 as I

    T = np.float32 if a.dtype == "f32" else np.float64
    w = 4 if T is np.float32 else 8
    x = tuple(np.linspace(0.0, 1.0, m + 1).astype(T) for m in a.grid)
    sg = I.Setup(x, {"u": ((I.PeriodicBC(), I.PeriodicBC()),) * 3})
    ps = I.psolver_spectral(sg)
    p = I.scalarfield(sg)
    p.normal_()
    cells = float(np.prod(a.grid))
    for mode in [0]:
        for _ in range(3):
            ps(p)
        torch.cuda.synchronize()
        ps.plan.profile_enable(True)
        ps.plan.profile_read()
        for _ in range(a.reps):
            ps(p)
        prof = ps.plan.profile_read()
        ps.plan.profile_enable(False)
        print(f"# grid {a.grid} {a.dtype}")
        for k in ("poisson", "fft_x_r2c", "fft_y_fwd", "fft_z_mid", "fft_y_inv", "fft_x_c2r"):
            ms = prof[k][0] / max(prof[k][1], 1)
            byt = (10 if k == "poisson" else 2) * w * cells
            print(f"{k:12s} {ms:8.3f} ms  {byt / ms / 1e6:8.0f} GB/s", flush=True)


if __name__ == "__main__":
    main()
