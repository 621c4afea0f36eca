#!/usr/bin/env python
"""bench.py -- headline benchmark of the insb200 hot path.

Metric (BASELINE.json): cell-updates/s per RK44 step on the TaylorGreenVortex3D workload,
reported with the fraction of the measured HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n 512] [--dtype f32|f64]
    python bench.py --impl reference ...     # CPU oracle ("port" of the reference schedule)

One "step" = one full RK44 time step (4 right-hand sides, 4 pressure projections) of the
whole grid.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the CPU arms below are meant to use every
# host core (the OpenMP runtime reads the variable when it is loaded, i.e. before the first oracle call)
if "TORCHELASTIC_RUN_ID" in os.environ and os.environ.get("OMP_NUM_THREADS") == "1":
    del os.environ["OMP_NUM_THREADS"]

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "incompressiblenavierstokes.jl_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

# ncu --set full, bytes per launch (average over the 4 launches of an RK44 step where they differ)
NCU_TRAFFIC_SOURCE = ("profiles/r3g_ncu_full_summary.md + profiles/r2v_ncu_full_summary.md (dram__bytes_read.sum + dram__bytes_write.sum per launch; "
                      "rhs_combine = mean of the four fused stage launches of a step: 5.95 (first) + 2 x 9.79 (middle) + 6.36 (last) GB)")
# fused periodic step (round 2); the stage-by-stage path of round 1 moved 6.58 / 2.15 / 3.73 / 5.27 GB
NCU_TRAFFIC = {"rhs_combine": 7.97e9, "divergence": 2.15e9, "project": 3.73e9, "poisson": 5.29e9}

METRIC = "cell-updates/s per RK44 step (TGV3D N^3)"
UNIT = "cell-updates/s"
NU = 1e-3


def tgv3d_u(dim, x, y, z):
    """examples/TaylorGreenVortex3D.jl:38-45"""
    if dim == 0:
        return np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y) * np.sin(2 * np.pi * z) / 2
    if dim == 1:
        return -np.cos(2 * np.pi * x) * np.sin(2 * np.pi * y) * np.sin(2 * np.pi * z) / 2
    return 0 * x + 0 * y + 0 * z


def fixed_dt(n):
    # 0.5 x propose_timestep at t = 0 (diffusive limit h^2/(6 nu) vs h/|u|max = 2h)
    h = 1.0 / n
    return 0.5 * min(h * h / NU / 6, 2 * h)


def weak_dims(n, world):
    d = [n, n, n]
    k, a = world, 2
    while k > 1:
        d[a] *= 2
        k //= 2
        a = (a - 1) % 3
    return tuple(d)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def ready(self):
        return self.proc is None or len(self.lines) > 0

    def mark(self):
        """start of the timed region: only samples taken from here on are reported"""
        self.t0 = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0 = getattr(self, "t0", 0.0)
        inside = [ln for ts, ln in self.lines if ts >= t0]
        if not inside and self.lines:  # region shorter than the sampling period: the sample closest to it
            inside = [self.lines[-1][1]]
        for ln in inside:
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------
# CPU arm: the oracle executing the reference's unfused schedule
# ----------------------------------------------------------------------------------------
def cpu_setup(n, dtype):
    import oracle as O
    from oracle import ckernels, initializers as oini, operators as oops, pressure as opr, timestep as ots

    cores = len(os.sched_getaffinity(0))
    opr.WORKERS = cores
    # stencil kernels: the C/OpenMP restatement on all host cores (the reference's CPU kernels are
    # threaded); broadcasts stay serial NumPy like Julia's; FFTs: scipy (pocketfft) on all cores
    oops.BACKEND = "c"
    ckernels.set_threads(cores)
    T = np.float32 if dtype == "f32" else np.float64
    x = tuple(np.linspace(0.0, 1.0, n + 1).astype(T) for _ in range(3))
    so = O.Setup(x, {"u": ((O.PeriodicBC(), O.PeriodicBC()),) * 3})
    ps = opr.psolver_spectral(so)
    u0 = oini.velocityfield(so, tgv3d_u, psolver=ps)
    m = ots.RK44(T)
    st = {"u": u0}
    cache = ots.get_cache(m, st, so)
    sp = ots.create_stepper(m, so, ps, st, 0.0)
    f = ots.make_navierstokes(so, NU)
    dt = T(fixed_dt(n))
    return dict(so=so, ps=ps, m=m, cache=cache, sp=sp, f=f, dt=dt, cores=cores, ots=ots, T=T)


def cpu_steps(n, dtype, steps, warmup, budget_s=None):
    """-> (seconds, steps executed, cores); with a time budget the loop stops early (never before one step)"""
    c = cpu_setup(n, dtype)
    ots, sp = c["ots"], c["sp"]
    for _ in range(warmup):
        sp = ots.timestep_(c["m"], c["f"], sp, c["dt"], c["cache"])
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        sp = ots.timestep_(c["m"], c["f"], sp, c["dt"], c["cache"])
        times.append(time.perf_counter() - t0)
        if budget_s is not None and sum(times) + times[-1] > budget_s:
            break
    assert np.all(np.isfinite(sp["state"]["u"]))
    return float(np.sum(times)), len(times), c["cores"]


CPU_SAMPLE = ("oracle executing the reference's unfused schedule (C/OpenMP stencil kernels on {cores} threads, "
              "scipy.fft workers={cores}, serial NumPy broadcasts)")


def run_reference(args):
    # the SAME configuration as the GPU arm (512^3), bounded in time instead of in size: up to
    # --steps steps, stopping once --ref-budget seconds of stepping are used up
    n = args.ref_n if args.ref_n > 0 else args.n
    warm = 1 if n <= 256 else 0
    tot, done, cores = cpu_steps(n, args.dtype, args.steps, warm, budget_s=args.ref_budget)
    val = n**3 * done / tot
    sample = (f"TGV3D {n}^3 {args.dtype} RK44, {done} of the {args.steps} requested steps (time budget {args.ref_budget:.0f} s), "
              f"{warm} warm-up, " + CPU_SAMPLE.format(cores=cores))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "steps_requested": args.steps, "warmup": warm, "ms_per_step": 1e3 * tot / done, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": (f"TaylorGreenVortex3D {args.n}^3 periodic {args.dtype} RK44 psolver_spectral" if args.gpus == 1 else
                                "DecayingTurbulence3D periodic %s RK44 psolver_spectral, weak-scaling series with %d^3 cells per GPU"
                                % (args.dtype, args.n)),
                   "sample": f"{n}^3 grid" + ("" if n == args.n else " (bounded CPU sample of the same workload)")
                   + ("" if args.gpus == 1 else "; one GPU's share of the box, stepped from the Taylor-Green field "
                      "(the cost per cell of the operator sequence does not depend on the field)")},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl")
    import insb200 as I

    T = np.float32 if args.dtype == "f32" else np.float64
    w = 4 if args.dtype == "f32" else 8
    n = args.n
    # weak-scaling series: n^3 cells per GPU; the grid doubles along z, then y, then x, so that
    # N = 1 is BASELINE config[1] (512^3) and N = 8 is config[2]'s 1024^3 box
    dims = weak_dims(n, world)
    x = tuple(np.linspace(0.0, 1.0, m + 1).astype(T) for m in dims)
    nid = I.share_unique_id() if world > 1 else None
    sg = I.Setup(x, {"u": ((I.PeriodicBC(), I.PeriodicBC()),) * 3}, rank=rank, world=world, nccl_id=nid)
    ps = I.psolver_spectral(sg)
    # N = 1: BASELINE config[1], the Taylor-Green vortex; N > 1: config[2], DecayingTurbulence3D -- the random_field
    # initial condition of examples/DecayingTurbulence3D.jl:31 (Orlandi spectrum, total energy 1), synthesised on the
    # slabs by the library (insb200_random_field, noise keyed by the global cell index)
    decaying = args.workload == "decaying" or (args.workload == "auto" and world > 1)
    if decaying:
        u = I.random_field(sg, 0.0, rng=20261018, psolver=ps)
        umax0 = I.maxabs(u, sg, plan=ps.plan)
    else:
        u = I.velocityfield(sg, tgv3d_u, psolver=ps)
        umax0 = 0.5
    method = I.RKMethods.RK44(T)
    params = dict(viscosity=NU)
    # eltype(u): the stage combine then stays in the field type
    dt = T(fixed_dt(max(dims)) if not decaying else 0.5 * min((1.0 / max(dims)) ** 2 / NU / 6, 1.0 / max(dims) / umax0))
    stepper = I.create_stepper(method, setup=sg, psolver=ps, state={"u": u}, t=0.0)
    ke0 = float(I.total_kinetic_energy(u, sg, plan=ps.plan))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # K steps = one solve_unsteady call with the time loop inside the library (insb200_solve_unsteady), like the
    # reference's users run it; --stepping calls makes one library call per step instead (the state is then projected
    # into the caller's array after every step instead of staying lazy between steps)
    def advance(stepper, k):
        if args.stepping == "calls":
            for _ in range(k):
                stepper = I.timestep_(method, I.navierstokes_, stepper, dt, params=params)
            return stepper
        t0 = stepper["t"]
        out, _ = I.solve_unsteady(setup=sg, tlims=(t0, t0 + k * float(dt)), start=stepper["state"], docopy=False,
                                  method=method, psolver=ps, dt=dt, params=params, native=True, graph=False)
        return I.create_stepper(method, setup=sg, psolver=ps, state={"u": out["u"]}, t=out["t"], n=stepper["n"] + k)

    stepper = advance(stepper, max(args.warmup, 3))
    barrier()
    ps.plan.profile_enable(True)
    ps.plan.profile_read()
    l0 = ps.plan.launch_count
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # keep stepping (untimed) until nvidia-smi delivers its first sample, so that the clocks line
    # really covers the timed region; all ranks take the same number of extra steps
    for _ in range(200):
        flag = torch.tensor([1 if (rank != 0 or sampler.ready()) else 0], device="cuda")
        if world > 1:
            dist.broadcast(flag, src=0)
        if int(flag.item()):
            break
        stepper = I.timestep_(method, I.navierstokes_, stepper, dt, params=params)
    ps.plan.profile_read()
    l0 = ps.plan.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark()
    e0.record()
    stepper = advance(stepper, args.steps)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = ps.plan.launch_count - l0
    prof = ps.plan.profile_read()
    ps.plan.profile_enable(False)
    if world > 1:
        tms = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    umax = I.maxabs(stepper["state"]["u"], sg, plan=ps.plan)
    assert np.isfinite(umax) and 0 < umax < 4 * umax0 + 1, f"solution blew up: max|u| = {umax}"
    # ---- self-check of the state the timed steps produced (a wrong halo / transpose that stays finite
    # would still break these): post-projection divergence at round-off level, and the kinetic energy
    # of the Taylor-Green vortex against its viscous decay KE(0) exp(-2 nu |k|^2 t), |k|^2 = 12 pi^2
    # (the energy-conserving convection moves energy between modes without dissipating it; at these
    # early times the decay is the linear one to O(1e-4))
    div = I.divergence_(I.scalarfield(sg), stepper["state"]["u"], sg)
    max_div = float(div[1:-1, 1:-1, 1:-1].abs().max().item())  # (diagnostic outside the timed region)
    del div
    if world > 1:
        td = torch.tensor([max_div], device="cuda", dtype=torch.float64)
        dist.all_reduce(td, op=dist.ReduceOp.MAX)
        max_div = float(td.item())
    ke = float(I.total_kinetic_energy(stepper["state"]["u"], sg, plan=ps.plan))
    t_end = float(stepper["t"])
    eps = float(np.finfo(T).eps)
    selfcheck = {"t": t_end, "max_abs_u": umax, "max_div_after_projection": max_div,
                 "max_div_bound": 50 * eps * umax * max(dims), "kinetic_energy": ke, "kinetic_energy_t0": ke0}
    assert max_div < selfcheck["max_div_bound"], f"projected field is not divergence free: {selfcheck}"
    if decaying:
        # no closed form: the energy-conserving scheme can only lose energy through viscosity; the synthesised
        # field starts at the prescribed total energy 1 (random_field) and decays slowly over these few steps
        selfcheck["kinetic_energy_target_t0"] = 1.0
        assert abs(ke0 - 1.0) < 1e-3, f"random_field missed its total energy: {selfcheck}"
        assert 0.9 * ke0 < ke < ke0, f"kinetic energy did not decay monotonically: {selfcheck}"
    else:
        ke_expected = ke0 * float(np.exp(-2.0 * NU * 12.0 * np.pi**2 * t_end))
        selfcheck.update(kinetic_energy_expected=ke_expected, kinetic_energy_rel_dev=abs(ke / ke_expected - 1.0))
        assert selfcheck["kinetic_energy_rel_dev"] < 1e-3, f"kinetic energy off the Taylor-Green decay: {selfcheck}"

    cells = int(np.prod(dims))  # one z-slab of the global grid per GPU
    value = cells * args.steps / (ms * 1e-3)
    # NVLink traffic per GPU per step: 2 all-to-all transposes per solve, 4 solves;
    # halo: u (3 comps) twice + p once per stage, 2 planes each
    nvlink = None
    if world > 1:
        m0p = (dims[0] // 2 + 1 + 7) // 8 * 8
        spec_local = m0p * dims[1] * (dims[2] // world) * 2 * w
        a2a = 8 * spec_local * (world - 1) / world
        halo = 4 * 7 * 2 * (dims[0] + 2) * (dims[1] + 2) * w
        sent = a2a + halo
        nvlink = {"bytes_sent_per_gpu_per_step": sent, "all_to_all_bytes": a2a, "halo_bytes": halo,
                  "achieved_GBps": sent / (ms / args.steps * 1e-3) / 1e9, "peak_GBps": 770.0,
                  "frac": sent / (ms / args.steps * 1e-3) / 1e9 / 770.0,
                  "peak_source": "measured peer copy per direction (B200_PROFILING.md)"}
    peak, peak_src = peaks()

    # ---- roofline of the dominant kernel group (DESIGN.md "Algorithmic bytes") ----
    # per launch and per cell: stage kernel 48w/4, divergence 4w, Poisson pass 2w (5 per solve),
    # projection 7w; bytes = count * w * n^3
    # with divergence! fused into the first FFT pass (no d array) that pass reads v (3w) and writes
    # the half spectrum (1w): Poisson 12w per solve, no divergence launches
    # fused periodic path (csrc/stage_fused.cu): one stencil launch per stage does the lazy projection of the
    # previous stage, the right-hand side, the stage combine and the divergence -- (10 + 17 + 17 + 11) w / 4
    # per launch -- and only the last stage of a step is projected explicitly
    fused_step = prof["divergence"][1] == 0 and 0 < prof["project"][1] < prof["rhs_combine"][1]
    fused_div = prof["divergence"][1] == 0 and not fused_step
    alg = {"rhs_combine": 13.75 if fused_step else 12.0, "divergence": 4.0, "poisson": 12.0 if fused_div else 10.0,
           "project": 7.0}
    dom = max(alg, key=lambda k: prof[k][0])
    alg.update({k: 2.0 for k in ("fft_x_r2c", "fft_y_fwd", "fft_z_mid", "fft_y_inv", "fft_x_c2r")})
    if fused_div:
        alg["fft_x_r2c"] = 4.0
    d_ms, d_cnt = prof[dom]
    per = d_ms / max(d_cnt, 1)
    achieved = alg[dom] * w * n**3 / (per * 1e-3) / 1e9
    groups = {k: {"ms_per_step": prof[k][0] / args.steps, "launch_groups": prof[k][1],
                  "GBps": (alg[k] * w * n**3 / (prof[k][0] / max(prof[k][1], 1) * 1e-3) / 1e9) if k in alg and prof[k][0] > 0 else None}
              for k in prof}
    step_bytes = 120.0 * w * n**3  # per GPU
    # dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the
    # committed `ncu --set full` capture of this workload (profiles/, 512^3 f32 only)
    traffic = NCU_TRAFFIC.get(dom) if (n == 512 and args.dtype == "f32") else None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": NCU_TRAFFIC_SOURCE if traffic else None,
                "peak_source": peak_src,
                "algorithmic_bytes_per_cell": alg[dom] * w,
                "step": {"achieved": step_bytes / (ms / args.steps * 1e-3) / 1e9,
                         "frac": step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak,
                         "algorithmic_bytes_per_cell_update": 120.0 * w,
                         # chained steps: first stage 14 w (it also stores the projected input), no projection pass
                         "moved_bytes_per_cell_update": ((99.0 if prof["project"][1] * 2 <= args.steps else 102.0)
                                                         if fused_step else 136.0) * w},
                "groups": groups}

    # ---- end to end through the C ABI with HOST buffers (pinned), every step H2D + D2H ----
    e2e = None
    if not args.no_e2e:  # every rank moves its own slab
        import ctypes as C

        from insb200._lib import lib, check
        from insb200.api import _params, _coef_dtype

        ke = max(2, min(args.steps, 3 if world > 1 else 4))  # pinned host states per rank (1.6 GB each at 512^3)
        uhs = [torch.empty(u.shape, dtype=u.dtype, pin_memory=True) for _ in range(ke)]
        for uh in uhs:
            uh.copy_(stepper["state"]["u"])
        uh = uhs[0]
        s = len(method.b)
        A = (C.c_double * (s * s))(*[float(v) for v in np.asarray(method.A).reshape(-1)])
        cc = (C.c_double * s)(*[float(v) for v in method.c])
        pr = _params(sg, viscosity=NU)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        # the ke pinned states form a ring: member m works on state m % ke, so a call with 3 ke members steps every
        # state three times, each time through host memory (the D2H of member m precedes the H2D of member m + ke on
        # the library's copy streams); the longer run amortises the fill and drain of the pipeline
        nmem = 3 * ke
        ptrs = (C.c_void_p * nmem)(*[uhs[m % ke].data_ptr() for m in range(nmem)])

        def host_steps(nm):
            # one call = nm independent pinned host states, each: H2D, one fused RK44 step, D2H;
            # the copies of neighbouring members overlap the kernels (insb200.h)
            check(lib().insb200_step_rk_host_ensemble(ps.plan.h, A, cc, s, _coef_dtype(method.T, dt), ptrs, None, nm,
                                                      0.0, float(dt), 1, C.byref(pr), st))

        def timed(fn):
            barrier()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            el = time.perf_counter() - t0
            if world > 1:
                tel = torch.tensor([el], device="cuda", dtype=torch.float64)
                dist.all_reduce(tel, op=dist.ReduceOp.MAX)
                el = float(tel.item())
            return el

        host_steps(3)
        el = timed(lambda: host_steps(nmem))

        # the same state stepped sequentially: every step waits for its own H2D and D2H (no overlap)
        def seq_steps():
            for _ in range(ke):
                check(lib().insb200_step_rk_host(ps.plan.h, A, cc, s, _coef_dtype(method.T, dt), C.c_void_p(uhs[0].data_ptr()), None, 0.0,
                                                 float(dt), 1, C.byref(pr), st))

        seq_steps()
        el_seq = timed(seq_steps)
        nbytes = uh.numel() * uh.element_size() * world
        e2e = {"value": cells * nmem / el, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
               "steps": nmem, "ms_per_step": 1e3 * el / nmem,
               "sequential": {"value": cells * ke / el_seq, "unit": UNIT, "ms_per_step": 1e3 * el_seq / ke,
                              "note": "insb200_step_rk_host on ONE state: H2D, fused RK44 step, D2H, synchronise, repeat"},
               "note": "insb200_step_rk_host_ensemble on every rank: each step = one pinned host state (slab of u) "
                       "-> device, one fused RK44 step, device -> host; the copies of the neighbouring steps overlap "
                       "the kernels of step i (3 streams, 3 device slots, a ring of %d pinned states stepped round-robin); "
                       "bytes summed over ranks, time = max over ranks" % ke}
        assert all(np.isfinite(float(uh_.abs().max())) for uh_ in uhs)
        del uhs

    # ---- parity at full size + CPU baseline beside it (N = 1 only) ----
    # rank 0 runs the oracle (the reference's unfused schedule on all host cores) on the SAME grid from the
    # same initial field: step 1 is compared with the CUDA path (north_star: rel L2 <= 1e-5 in Float32,
    # 1e-12 in Float64), step 2 is the timed CPU sample.
    cpu = None
    parity = None
    if world == 1 and not args.no_cpu:
        nref = args.ref_n if args.ref_n > 0 else n
        c = cpu_setup(nref, args.dtype)
        ots = c["ots"]
        do_check = nref == n and not args.no_check
        if do_check:
            ug = I.to_device(sg, c["sp"]["state"]["u"])
            spg = I.create_stepper(method, setup=sg, psolver=ps, state={"u": ug}, t=0.0)
        sp = ots.timestep_(c["m"], c["f"], c["sp"], c["dt"], c["cache"])
        if do_check:
            spg = I.timestep_(method, I.navierstokes_, spg, dt, params=params)
            got = spg["state"]["u"].cpu().numpy().T
            ref = sp["state"]["u"]
            num = float(np.sqrt(np.sum((got.astype(np.float64) - ref) ** 2)))
            den = float(np.sqrt(np.sum(ref.astype(np.float64) ** 2)))
            dv = I.divergence_(I.scalarfield(sg), spg["state"]["u"], sg)
            parity = {"grid": [nref] * 3, "steps": 1, "rel_l2": num / den, "tol": 1e-5 if args.dtype == "f32" else 1e-12,
                      "max_div": float(dv[1:-1, 1:-1, 1:-1].abs().max().item()),
                      "against": "CPU oracle (oracle/, C/OpenMP + NumPy restatement of the reference), same initial field"}
            del got, ref, dv, ug, spg
            assert parity["rel_l2"] < parity["tol"], f"full-size parity failed: {parity}"
        t0 = time.perf_counter()
        sp = ots.timestep_(c["m"], c["f"], sp, c["dt"], c["cache"])
        tot = time.perf_counter() - t0
        cpu = {"value": nref**3 / tot, "unit": UNIT, "cores": c["cores"], "kind": "port",
               "sample": f"TGV3D {nref}^3 {args.dtype} RK44, 1 step after 1 warm-up step, " + CPU_SAMPLE.format(cores=c["cores"])}
        del c, sp

    # ---- P GPUs == 1 GPU on a small grid of the same problem class (N > 1) ----
    equality = None
    if world > 1 and not args.no_check:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import mgpu_worker

        ea = argparse.Namespace(grid=[128, 64, 128], dtype=args.dtype, steps=2)
        ok, rows = mgpu_worker.run_case(ea, "periodic", rank, world, quiet=True)
        if rank == 0:
            equality = {"grid": ea.grid, "steps": ea.steps, "ok": ok,
                        "max_rel_err": max(r["err"] for r in rows if r["case"] in ("project", "psolve", "RK44", "LMWray3")),
                        "cases": {r["case"]: r["err"] for r in rows},
                        "against": f"the same problem on ONE GPU (rank 0), {world}-GPU slab decomposition gathered"}
            assert ok, f"{world}-GPU result differs from the 1-GPU result: {rows}"

    if rank == 0:
        name = "DecayingTurbulence3D" if decaying else "TaylorGreenVortex3D"
        wl = f"{name} {dims[0]}x{dims[1]}x{dims[2]} periodic {args.dtype} RK44 psolver_spectral"
        if world > 1:
            wl += (f" (weak-scaling series: {n}^3 cells per GPU, BASELINE config[2] at 8 GPUs"
                   + (", random_field initial condition synthesised on the slabs)" if decaying else ")"))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": wl,
                       "n": list(dims), "cells_per_gpu": n**3, "nu": NU, "dt": float(dt), "method": f"RK44(T={args.dtype})",
                       "stepping": ("insb200_solve_unsteady: one call for the timed steps, the state stays lazy (unprojected "
                                    "velocity + pressure) between steps" if args.stepping == "loop" else
                                    "insb200_step_rk: one call per step"),
                       "l2": f"working set {14 * w * (n + 2)**3 / 1e9:.1f} GB per GPU >> 126 MB L2 (no flush needed)",
                       "parallelism": "1 GPU" if world == 1 else
                       f"z-slab decomposition over {world} GPUs (halo exchange + FFT transposes by peer stores over NVLink)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "selfcheck": selfcheck, "parity": parity,
        }
        if equality:
            line["equality_vs_1gpu"] = equality
        if nvlink:
            line["nvlink"] = nvlink
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="insb200", choices=["insb200", "reference"])
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--ref-n", type=int, default=0, help="grid of the CPU arms (0 = the workload's own grid)")
    ap.add_argument("--ref-budget", type=float, default=150.0, help="seconds of CPU stepping of --impl reference")
    ap.add_argument("--no-check", action="store_true", help="skip the full-size parity leg against the oracle")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--workload", default="auto", choices=["auto", "tgv", "decaying"],
                    help="auto: Taylor-Green vortex on one GPU (config[1]), decaying turbulence on several (config[2])")
    ap.add_argument("--stepping", default="loop", choices=["loop", "calls"],
                    help="loop: the timed steps are one insb200_solve_unsteady call; calls: one insb200_step_rk call per step")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) == 0:
            run_reference(args)
        return
    run_gpu(args)


if __name__ == "__main__":
    main()
