"""Worker of the multi-GPU equality tests ("P-GPU result == 1-GPU result").

Run under torchrun, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/mgpu_worker.py [--grid 64 32 64] [--dtype f64]

Every rank builds the same seeded global field, keeps its z-slab, and runs projection,
RK44 / LMWray3 steps, the CFL and max-abs reductions through the slab-decomposed plan;
rank 0 repeats everything with a world == 1 plan on its own GPU and compares.
Prints one JSON line per case on rank 0; exit code 1 on any mismatch.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "incompressiblenavierstokes.jl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, nargs=3, default=[64, 32, 64])
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--strict", action="store_true",
                    help="projection, Poisson solve and time steps must reproduce the 1-GPU result BIT FOR BIT (the kernels "
                         "of a slab and of the whole grid round alike); without it the north-star tolerances apply")
    ap.add_argument("--case", default="periodic", choices=["periodic", "channel", "rb", "both", "all", "timeout"],
                    help="periodic: psolver_spectral; channel: no-slip walls in y on a tanh-stretched grid, "
                         "psolver_tridiagonal (BASELINE config 5 shape); rb: Rayleigh-Benard box, walls in y AND in "
                         "the decomposed z, temperature, boussinesq!, psolver_transform (BASELINE config 4 shape)")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if os.environ.get("INSB200_HANG_DUMP"):  # debugging aid: Python stacks of all threads after N seconds
        import faulthandler

        faulthandler.dump_traceback_later(int(os.environ["INSB200_HANG_DUMP"]), exit=True)
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl")
    if args.case == "timeout":
        sys.exit(run_timeout(rank, world))
    ok = True
    cases = {"both": ["periodic", "channel"], "all": ["periodic", "channel", "rb"]}.get(args.case, [args.case])
    for case in cases:
        ok = run_case(args, case, rank, world)[0] and ok
        dist.barrier()
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


def run_timeout(rank, world):
    """A peer that never arrives must produce an error, not a hang and not stale data: rank 0 skips one halo
    exchange; every other rank gives up after INSB200_PEER_TIMEOUT_MS (the waiting kernel traps: sticky launch
    error) and must see an exception.  Exit code 0 = the ranks that waited raised."""
    import insb200 as I

    os.environ.setdefault("INSB200_PEER_TIMEOUT_MS", "1500")
    n = (64, 32, 64)
    x = tuple(np.linspace(0.0, 1.0, m + 1) for m in n)
    nid = I.share_unique_id()
    sd = I.Setup(x, {"u": ((I.PeriodicBC(), I.PeriodicBC()),) * 3}, rank=rank, world=world, nccl_id=nid)
    psd = I.psolver_spectral(sd)
    u = I.vectorfield(sd)
    I.apply_bc_u_(u, 0.0, sd, plan=psd.plan)  # one complete exchange: everybody takes part
    torch.cuda.synchronize()
    dist.barrier()
    raised = False
    if rank != 0:
        try:
            I.apply_bc_u_(u, 0.0, sd, plan=psd.plan)  # rank 0 never sends this one
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001  (torch reports the sticky launch failure)
            raised = True
            print(json.dumps({"case": "peer_timeout", "rank": rank, "error": str(e)[:120]}), flush=True)
    else:
        import time

        time.sleep(4.0)
    # no collective after this point: the contexts of the waiting ranks are poisoned by design
    os._exit(0 if (rank == 0 or raised) else 1)


def run_case(args, case, rank, world, quiet=False):
    import insb200 as I

    T = np.float64 if args.dtype == "f64" else np.float32
    tol = 1e-12 if T is np.float64 else 1e-5
    n = args.grid
    x = tuple(np.linspace(0.0, 1.0, m + 1).astype(T) for m in n)
    bcs = {"u": ((I.PeriodicBC(), I.PeriodicBC()),) * 3}
    mk = I.psolver_spectral
    if case == "channel":
        import oracle as O

        x = (np.linspace(0.0, 4.0, n[0] + 1).astype(T), O.tanh_grid(0, 2, n[1], 1.4).astype(T),
             np.linspace(0.0, 2.0, n[2] + 1).astype(T))
        bcs = {"u": ((I.PeriodicBC(), I.PeriodicBC()), (I.DirichletBC(), I.DirichletBC()),
                     (I.PeriodicBC(), I.PeriodicBC()))}
        mk = I.psolver_tridiagonal
    if case == "rb":  # examples/RayleighBenard3D.jl:23-38
        x = (np.linspace(0.0, np.pi, n[0] + 1).astype(T), np.linspace(0.0, 1.0, n[1] + 1).astype(T),
             np.linspace(0.0, 1.0, n[2] + 1).astype(T))
        bcs = {"u": ((I.PeriodicBC(), I.PeriodicBC()), (I.DirichletBC(), I.DirichletBC()),
                     (I.DirichletBC(), I.DirichletBC())),
               "temp": ((I.PeriodicBC(), I.PeriodicBC()), (I.SymmetricBC(), I.SymmetricBC()),
                        (I.DirichletBC(1.0), I.DirichletBC(0.0)))}
        mk = I.psolver_transform
    nid = I.share_unique_id()
    sd = I.Setup(x, bcs, rank=rank, world=world, nccl_id=nid)
    s1 = I.Setup(x, bcs) if rank == 0 else None
    psd = mk(sd)
    ps1 = mk(s1) if rank == 0 else None

    rng = np.random.default_rng(7)
    ug = np.asfortranarray(rng.standard_normal(sd.N + (3,)).astype(T))
    pg = np.asfortranarray(rng.standard_normal(sd.N).astype(T))
    ok = True
    out = []

    exact = {"project", "psolve", "RK44", "LMWray3", "RK44_temp", "LMWray3_temp", "LES_WALE_RK44", "vorticity",
             "RK44_native_loop", "LMWray3_native_loop"}

    def report(name, err, lim):
        nonlocal ok
        if getattr(args, "strict", False) and name in exact:
            lim = 0.0
        good = bool(err <= lim)
        ok = ok and good
        out.append({"case": name, "err": err, "tol": lim, "ok": good})

    # ---- ghost fills + projection -------------------------------------------------------
    ud = I.to_device(sd, ug)
    I.apply_bc_u_(ud, 0.0, sd, plan=psd.plan)
    pd = I.scalarfield(sd)
    I.project_(ud, sd, psolver=psd, p=pd)
    I.apply_bc_u_(ud, 0.0, sd, plan=psd.plan)
    div_d = I.divergence_(I.scalarfield(sd), ud, sd)
    got_u = I.gather_global(sd, ud)
    got_div = I.gather_global(sd, div_d)
    # ---- Poisson solve alone ------------------------------------------------------------
    pd2 = I.to_device(sd, pg)
    psd(pd2)
    I.apply_bc_p_(pd2, 0.0, sd, plan=psd.plan)
    got_p = I.gather_global(sd, pd2)
    # ---- time stepping ------------------------------------------------------------------
    dt = T(2e-3)
    params = dict(viscosity=2e-3)
    force = I.navierstokes_
    tg = None
    if case == "rb":
        params = dict(viscosity=2e-3, conductivity=2e-3, gravity=1.0, gdir=2, dodissipation=True)
        force = I.boussinesq_
        tg = np.asfortranarray(0.5 + 0.1 * rng.standard_normal(sd.N).astype(T))

    def start_state(setup, u):
        st = {"u": u.clone()}
        if tg is not None:
            st["temp"] = I.apply_bc_temp_(I.to_device(setup, tg), 0.0, setup)
        return st

    res = {}
    for name, method in (("RK44", I.RKMethods.RK44(T)), ("LMWray3", I.LMWray3(T))):
        st = I.create_stepper(method, setup=sd, psolver=psd, state=start_state(sd, ud), t=0.0)
        for _ in range(args.steps):
            st = I.timestep_(method, force, st, dt, params=params)
        res[name] = I.gather_global(sd, st["state"]["u"])
        if tg is not None:
            res[name + "_temp"] = I.gather_global(sd, st["state"]["temp"])
        res[name + "_cfl"] = float(I.propose_timestep(force, st["state"], sd, params, plan=psd.plan))
        res[name + "_umax"] = I.maxabs(st["state"]["u"], sd, plan=psd.plan)
    # eddy-viscosity closure inside the stepper (9 tensor components through the halo exchange),
    # total kinetic energy (all-reduced sum), vorticity
    les = dict(viscosity=2e-3, eddyviscosity=I.WALE(0.569))
    st = I.create_stepper(I.RKMethods.RK44(T), setup=sd, psolver=psd, state={"u": ud.clone()}, t=0.0)
    for _ in range(args.steps):  # (also with walls in the decomposed direction: zero_out_wall! on the end ranks)
        st = I.timestep_(I.RKMethods.RK44(T), I.navierstokes_, st, dt, params=les)
    res["LES"] = I.gather_global(sd, st["state"]["u"])
    res["tke"] = float(I.total_kinetic_energy(st["state"]["u"], sd, plan=psd.plan))
    res["vort"] = I.gather_global(sd, I.vorticity(st["state"]["u"], sd))

    # the time loop inside the library (chained steps on the fused periodic path: the state stays lazy between steps)
    if case == "periodic":
        for name, method in (("RK44", I.RKMethods.RK44(T)), ("LMWray3", I.LMWray3(T))):
            fin, _ = I.solve_unsteady(setup=sd, tlims=(0.0, 3 * float(dt)), start={"u": ud}, method=method, psolver=psd,
                                      dt=dt, params=params, native=True)
            res[name + "_native_loop"] = I.gather_global(sd, fin["u"])

    # random_field synthesised on the slabs (distributed transforms, all-reduced shell sums): injected host noise
    # and the device generator, which is keyed by the GLOBAL cell index
    if case == "periodic":
        for tag, r in (("host_noise", np.random.default_rng(11)), ("device_noise", 1234)):
            res["rf_" + tag] = I.gather_global(sd, I.random_field(sd, 0.0, totalenergy=0.5, kpeak=4, rng=r, psolver=psd))

    if rank == 0:
        if case == "periodic":
            u1s = I.to_device(s1, ug)
            I.apply_bc_u_(u1s, 0.0, s1)
            I.project_(u1s, s1, psolver=ps1, p=I.scalarfield(s1))
            I.apply_bc_u_(u1s, 0.0, s1)
            for name, method in (("RK44", I.RKMethods.RK44(T)), ("LMWray3", I.LMWray3(T))):
                # one GPU, the same loop with chaining switched off: u is materialised after every step (that the
                # chained loop equals call-by-call stepping bit for bit is tests/test_gpu_variants.py's business)
                os.environ["INSB200_NO_CHAINED_STEPS"] = "1"
                try:
                    ref, _ = I.solve_unsteady(setup=s1, tlims=(0.0, 3 * float(dt)), start={"u": u1s}, method=method,
                                              psolver=ps1, dt=dt, params=params, native=True, graph=False)
                finally:
                    del os.environ["INSB200_NO_CHAINED_STEPS"]
                report(name + "_native_loop", rel(res[name + "_native_loop"], I.to_host(ref["u"])), tol)
            for tag, r in (("host_noise", np.random.default_rng(11)), ("device_noise", 1234)):
                f1 = I.to_host(I.random_field(s1, 0.0, totalenergy=0.5, kpeak=4, rng=r, psolver=ps1))
                report("random_field_" + tag, rel(res["rf_" + tag], f1), 1e-13 if T is np.float64 else 2e-6)
        u1 = I.to_device(s1, ug)
        I.apply_bc_u_(u1, 0.0, s1)
        p1 = I.scalarfield(s1)
        I.project_(u1, s1, psolver=ps1, p=p1)
        I.apply_bc_u_(u1, 0.0, s1)
        report("project", rel(got_u, I.to_host(u1)), tol)
        inner = (slice(1, -1),) * 3
        # (a property of the projected N(0, 1) field, not an equality: round-off of the solve scales with the extents)
        report("divergence_after_project", float(np.abs(got_div[inner]).max()),
               1e-10 if T is np.float64 else 2e-5 * max(n) * (4 if case == "rb" else 1))
        p1b = I.to_device(s1, pg)
        ps1(p1b)
        I.apply_bc_p_(p1b, 0.0, s1)
        report("psolve", rel(got_p, I.to_host(p1b)), tol)
        for name, method in (("RK44", I.RKMethods.RK44(T)), ("LMWray3", I.LMWray3(T))):
            st = I.create_stepper(method, setup=s1, psolver=ps1, state=start_state(s1, u1), t=0.0)
            for _ in range(args.steps):
                st = I.timestep_(method, force, st, dt, params=params)
            report(name, rel(res[name], I.to_host(st["state"]["u"])), tol)
            if tg is not None:
                report(name + "_temp", rel(res[name + "_temp"], I.to_host(st["state"]["temp"])), tol)
            c1 = float(I.propose_timestep(force, st["state"], s1, params, plan=ps1.plan))
            report(name + "_cfl", abs(res[name + "_cfl"] - c1) / c1, 1e-6)
            m1 = I.maxabs(st["state"]["u"], s1)
            report(name + "_umax", abs(res[name + "_umax"] - m1) / m1, 1e-6)
        st = I.create_stepper(I.RKMethods.RK44(T), setup=s1, psolver=ps1, state={"u": u1.clone()}, t=0.0)
        for _ in range(args.steps):
            st = I.timestep_(I.RKMethods.RK44(T), I.navierstokes_, st, dt, params=les)
        report("LES_WALE_RK44", rel(res["LES"], I.to_host(st["state"]["u"])), tol)
        e1 = float(I.total_kinetic_energy(st["state"]["u"], s1))
        report("total_kinetic_energy", abs(res["tke"] - e1) / e1, 1e-12 if T is np.float64 else 1e-6)
        w1 = I.to_host(I.vorticity(st["state"]["u"], s1))
        inner = (slice(1, -1),) * 3
        report("vorticity", rel(res["vort"][inner], w1[inner]), tol)
        for o in out:
            o.update(world=world, n=n, dtype=args.dtype, problem=case)
            if not quiet:
                print(json.dumps(o), flush=True)
    # collective tear-down (peer mappings closed on every rank before anybody frees exported memory)
    I.release_plans(sd)
    return ok, out


if __name__ == "__main__":
    main()
