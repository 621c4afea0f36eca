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
    ap.add_argument("--no-cufft", action="store_true", help="skip the cuFFT comparison leg")
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
    if not a.no_cufft:
        del p
        cufft_leg(a.grid, T, a.reps)


def cufft_leg(grid, T, reps):
    """Comparison only (never on the product path): the same solve with the vendor FFT -- cuFFT 3-D R2C / C2R
    through torch.fft plus the elementwise eigenvalue divide (SURVEY K7's "bar to beat"), timed with CUDA
    events on the same box.  The interior copy in / out of the ghosted array is included, like view(p, Ip)."""
    n = grid
    dev = "cuda"
    rt = torch.float32 if T is np.float32 else torch.float64
    p = torch.randn((n[2] + 2, n[1] + 2, n[0] + 2), device=dev, dtype=rt)
    ax = [torch.arange(m, device=dev, dtype=torch.float64) for m in n]
    # eigenvalues of the periodic second difference (pressure.jl:336-342), Float64 then rounded
    e = [(4 * (m * m) * torch.sin(torch.pi * a_ / m) ** 2) for a_, m in zip(ax, n)]
    e0 = e[0][: n[0] // 2 + 1]
    den = (e0[None, None, :] + e[1][None, :, None] + e[2][:, None, None]).to(rt)
    den[0, 0, 0] = 1
    inv = (-1.0 / den).contiguous()
    inv[0, 0, 0] = 0

    def solve():
        v = p[1:-1, 1:-1, 1:-1]
        ph = torch.fft.rfftn(v, dim=(0, 1, 2))
        ph *= inv
        v.copy_(torch.fft.irfftn(ph, s=(n[2], n[1], n[0]), dim=(0, 1, 2)))

    for _ in range(3):
        solve()
    torch.cuda.synchronize()
    a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a_.record()
    for _ in range(reps):
        solve()
    b_.record()
    torch.cuda.synchronize()
    ms = a_.elapsed_time(b_) / reps
    w = 4 if T is np.float32 else 8
    print(f"cuFFT leg (torch.fft.rfftn + divide + irfftn, comparison only): {ms:8.3f} ms per solve  "
          f"{10 * w * float(np.prod(n)) / ms / 1e6:8.0f} GB/s of the 10 w algorithmic bytes", flush=True)


if __name__ == "__main__":
    main()
