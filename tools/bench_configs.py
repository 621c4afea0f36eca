#!/usr/bin/env python
"""Times the other BASELINE.json configurations on ONE GPU (they are parity-test cases, not
bench.py lines): cfg1 TaylorGreenVortex2D 256^2 F64, cfg4 RayleighBenard3D 512x512x256
(Boussinesq, psolver_transform), cfg5 ChannelFlow3D 1024x512x512 F64 (psolver_tridiagonal).
Prints one JSON line per case with ms/step, cell-updates/s and the per-group device times.

    python tools/bench_configs.py [--cases cfg1 cfg4 cfg5] [--scale 1.0]
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


def run(name, sg, ps, state, method, force, params, dt, steps, warm=2):
    import insb200 as I

    st = I.create_stepper(method, setup=sg, psolver=ps, state=state, t=0.0)
    for _ in range(warm):
        st = I.timestep_(method, force, st, dt, params=params)
    torch.cuda.synchronize()
    ps.plan.profile_enable(True)
    ps.plan.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        st = I.timestep_(method, force, st, dt, params=params)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    prof = ps.plan.profile_read()
    ps.plan.profile_enable(False)
    umax = I.maxabs(st["state"]["u"], sg, plan=ps.plan)
    cells = float(np.prod(sg.Np))
    if sg.rank == 0:
        print(json.dumps({"case": name, "N": list(sg.N), "dtype": sg.dtype.name, "ms_per_step": ms,
                      "cell_updates_per_s": cells / ms * 1e3, "umax": umax,
                      "groups_ms_per_step": {k: round(v[0] / steps, 3) for k, v in prof.items() if v[1]}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", nargs="+", default=["cfg1", "cfg4", "cfg5"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink every extent (tests)")
    ap.add_argument("--world", type=int, default=1, help="cfg4 / cfg5: z-slab decomposition (run under torchrun)")
    a = ap.parse_args()
    import insb200 as I

    rank, nid = 0, None
    if a.world > 1:
        import torch.distributed as dist

        rank = int(os.environ["RANK"])
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl")
        nid = I.share_unique_id()

    def sz(n):
        return max(8, int(round(n * a.scale / 2)) * 2)

    P, D, S = I.PeriodicBC, I.DirichletBC, I.SymmetricBC
    if "cfg1" in a.cases:  # examples/TaylorGreenVortex2D.jl:59-72
        n = sz(256)
        x = (np.linspace(0, 2 * np.pi, n + 1),) * 2
        sg = I.Setup(x, {"u": ((P(), P()),) * 2})
        ps = I.psolver_spectral(sg)
        u = I.velocityfield(sg, lambda d, x, y: (-np.sin(x) * np.cos(y) if d == 0 else np.cos(x) * np.sin(y)) + 0 * x * y,
                            psolver=ps)
        run("cfg1 TGV2D RK44", sg, ps, {"u": u}, I.RKMethods.RK44(), I.navierstokes_, dict(viscosity=5e-4), 0.02, 100)
        # the whole 100-step run through solve_unsteady: Python loop, native loop, native loop + CUDA graph
        import time

        for tag, opts in (("python loop", dict(native=False)), ("native loop", dict(native=True, graph=False)),
                          ("native loop + CUDA graph", dict(native=True, graph=True))):
            best = 1e9
            for _ in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                out, _ = I.solve_unsteady(setup=sg, tlims=(0.0, 2.0), start={"u": u}, method=I.RKMethods.RK44(),
                                          psolver=ps, dt=0.02, params=dict(viscosity=5e-4), **opts)
                torch.cuda.synchronize()
                best = min(best, time.perf_counter() - t0)
            print(json.dumps({"case": f"cfg1 TGV2D RK44 solve_unsteady 100 steps, {tag}", "wall_ms_per_step": best * 10,
                              "cell_updates_per_s": n * n * 100 / best}), flush=True)
    if "cfg4" in a.cases:  # examples/RayleighBenard3D.jl:23-79
        for T in (np.float64, np.float32):
            n = (sz(512), sz(512), sz(256))
            x = (np.linspace(0, np.pi, n[0] + 1).astype(T), np.linspace(0, 1, n[1] + 1).astype(T),
                 np.linspace(0, 1, n[2] + 1).astype(T))
            nid = I.share_unique_id() if a.world > 1 else None  # (an NCCL id serves one communicator)
            sg = I.Setup(x, {"u": ((P(), P()), (D(), D()), (D(), D())),
                             "temp": ((P(), P()), (S(), S()), (D(1.0), D(0.0)))}, rank=rank, world=a.world, nccl_id=nid)
            ps = I.psolver_transform(sg)
            u = I.vectorfield(sg)
            temp = I.temperaturefield(sg, lambda x, y, z: 0.5 + np.sin(20 * x) * np.sin(np.pi * 20 * y) / 100 + 0 * z)
            params = dict(viscosity=2.5e-4, conductivity=2.5e-4, gravity=1.0, gdir=2, dodissipation=True)
            for mname, m in (("RK44", I.RKMethods.RK44(T)), ("LMWray3", I.LMWray3(T))):
                run(f"cfg4 RayleighBenard3D {mname} {np.dtype(T).name} ({a.world} GPU)", sg, ps,
                    {"u": u.clone(), "temp": temp.clone()}, m, I.boussinesq_, params, T(1e-3), 5)
            I.release_plans(sg)
            del sg, ps, u, temp
            torch.cuda.empty_cache()
    if "cfg5" in a.cases:  # examples/ChannelFlow.jl:33-167
        n = (sz(1024), sz(512), sz(512))
        x = (np.linspace(0, 4 * np.pi, n[0] + 1), I.tanh_grid(0.0, 2.0, n[1], 1.4), np.linspace(0, 4 * np.pi / 3, n[2] + 1))
        nid = I.share_unique_id() if a.world > 1 else None
        sg = I.Setup(x, {"u": ((P(), P()), (D(), D()), (P(), P()))}, rank=rank, world=a.world, nccl_id=nid)
        ps = I.default_psolver(sg)
        assert ps.kind == "tridiagonal"

        def U(d, x, y, z):
            if d == 0:
                return (1 - (y - 1) ** 2) * (1 + 0.1 * np.sin(3 * z) * np.cos(2 * x))
            return 0.05 * np.sin(2 * x) * np.sin(np.pi * y) * np.cos(3 * z) + 0 * (x + y + z)

        u = I.velocityfield(sg, U, psolver=ps)
        if a.world == 1:  # the example's LES variant: navierstokes! + eddy_viscosity_closure!(WALE) (ChannelFlow.jl:177-182)
            run("cfg5 ChannelFlow3D RK44 + WALE closure (1 GPU)", sg, ps, {"u": u.clone()}, I.RKMethods.RK44(), I.navierstokes_,
                dict(viscosity=1 / 180, bodyforce=(1.0, 0.0, 0.0), eddyviscosity=I.WALE(0.569)), 1e-4, 3)
        run(f"cfg5 ChannelFlow3D RK44 ({a.world} GPU)", sg, ps, {"u": u}, I.RKMethods.RK44(), I.navierstokes_,
            dict(viscosity=1 / 180, bodyforce=(1.0, 0.0, 0.0)), 1e-4, 3)


if __name__ == "__main__":
    main()
