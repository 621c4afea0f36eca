"""Stage-by-stage check of the fused periodic path (csrc/stage_fused.cu) against the oracle.

Runs the fused RK44 step with INSB200_FUSED_STOP_AFTER = 1..4, copies the padded internal arrays back
(insb200_debug_fused_copy) and compares their interior -- and their ghost layers -- with the oracle's
v_i (unprojected stage velocity), acc_i, d_i (scaled divergence) and p_i.

    python tools/debug_fused.py [n0 n1 n2] [f32|f64]
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "incompressiblenavierstokes.jl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402

import oracle as O  # noqa: E402
from oracle import bc as obc, operators as oops, pressure as opr, timestep as ots  # noqa: E402
from gpu_helpers import make_setups, smooth_vec, uniform  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def fetch(I, plan, which, T):
    from insb200._lib import lib, check

    dims = (C.c_int * 4)()
    check(lib().insb200_debug_fused_copy(plan.h, which, None, 0, dims))
    npx, npy, npz, nc = list(dims)
    buf = np.empty(npx * npy * npz * nc, dtype=T)
    check(lib().insb200_debug_fused_copy(plan.h, which, buf.ctypes.data, buf.nbytes, dims))
    a = buf.reshape((nc, npz, npy, npx)).transpose(3, 2, 1, 0)  # (x, y, z, c)
    return a if nc == 3 else a[..., 0]


def padded_from_ref(ref, n, ghost=2):
    """periodic extension of a reference-layout array (1 ghost) to `ghost` layers, as the internal layout
    stores it: index (i + 3, j + 1, k + 1) for 1-based cell (i, j, k)"""
    inner = ref[1:-1, 1:-1, 1:-1]
    idx = [np.arange(-ghost + 1, m + ghost + 1 - 0) for m in n]  # cells 1-ghost .. n+ghost (1-based)
    w = [((ix - 1) % m) for ix, m in zip(idx, n)]
    return inner[np.ix_(w[0], w[1], w[2])]


def main():
    import insb200 as I
    import torch

    args = sys.argv[1:]
    T = np.float64
    if args and args[-1] in ("f32", "f64"):
        T = np.float32 if args[-1] == "f32" else np.float64
        args = args[:-1]
    n = tuple(int(v) for v in args[:3]) if len(args) >= 3 else (32, 64, 32)
    so, sg = make_setups((uniform(n[0]), uniform(n[1], 0.5), uniform(n[2], 2.0)), ("P", "P", "P"), T=T)
    ps_o, ps_g = opr.psolver_spectral(so), I.psolver_spectral(sg)
    u0 = opr.project(smooth_vec(so, 45), so, ps_o)
    obc.apply_bc_u(u0, 0.0, so)
    nu, dt = 2e-3, T(2e-3)
    m_o, m_g = ots.RK44(T), I.RKMethods.RK44(T)
    A, b = np.asarray(m_o.A), np.asarray(m_o.b)
    s = len(b)
    # ---- oracle, stage by stage (step_explicit_runge_kutta.jl:24-46) ----
    u = u0.copy(order="F")
    ks = []
    exp = []
    for i in range(s):
        k = {"u": O.vectorfield(so)}
        oops.navierstokes_(k, {"u": u}, 0.0, so, nu)
        ks.append(k["u"].copy(order="F"))
        v = u0.copy(order="F")
        for j in range(i + 1):
            v = (v + (T(dt) * T(A[i, j])) * ks[j]).astype(T)
        acc = u0.copy(order="F")
        for j in range(i + 1):
            acc = (acc + (T(dt) * T(A[s - 1, j])) * ks[j]).astype(T)
        v = np.asfortranarray(v)
        obc.apply_bc_u(v, 0.0, so)
        d = oops.divergence_(O.scalarfield(so), v, so)
        oops.scalewithvolume_(d, so)
        p = d.copy(order="F")
        ps_o(p)
        obc.apply_bc_p(p, 0.0, so)
        exp.append(dict(v=v.copy(), acc=acc, d=d, p=p.copy()))
        u = oops.applypressure_(v, p, so)
        obc.apply_bc_u(u, 0.0, so)
    L = lambda a: a  # noqa: E731
    for stop in range(1, s + 1):
        os.environ["INSB200_FUSED_STOP_AFTER"] = str(stop)
        sp = I.create_stepper(m_g, setup=sg, psolver=ps_g, state={"u": I.to_device(sg, u0)}, t=0.0)
        I.timestep_(m_g, I.navierstokes_, sp, dt, params=dict(viscosity=nu))
        torch.cuda.synchronize()
        e = exp[stop - 1]
        fv = fetch(I, ps_g.plan, (stop - 1) & 1, T)
        fd = fetch(I, ps_g.plan, 4, T)
        fp = fetch(I, ps_g.plan, 3, T)
        facc = fetch(I, ps_g.plan, 2, T)
        sl = (slice(4, 4 + n[0]), slice(2, 2 + n[1]), slice(2, 2 + n[2]))
        gl = (slice(2, 6 + n[0]), slice(0, 4 + n[1]), slice(0, 4 + n[2]))  # cells -1 .. n + 2
        inner = (slice(1, -1),) * 3
        out = [f"stage {stop}:"]
        for c in range(3):
            out.append(f"v{c} int {rel(fv[sl + (c,)], e['v'][inner + (c,)]):.2e} ghost {rel(fv[gl + (c,)], padded_from_ref(e['v'][..., c], n)):.2e}")
        if stop < s:
            for c in range(3):
                out.append(f"acc{c} {rel(facc[sl + (c,)], e['acc'][inner + (c,)]):.2e}")
        out.append(f"d {rel(fd[sl], e['d'][inner]):.2e}")
        out.append(f"p int {rel(fp[sl], e['p'][inner]):.2e} ghost {rel(fp[gl], padded_from_ref(e['p'], n)):.2e}")
        print("  ".join(out), flush=True)
        if stop == 1:  # where do the first-stage errors sit?
            for name, got, want in (("v0", fv[sl + (0,)], e["v"][inner + (0,)]), ("d", fd[sl], e["d"][inner])):
                diff = np.abs(got.astype(np.float64) - want)
                bad = np.argwhere(diff > 1e-6 * np.abs(want).max())
                print(f"   {name}: {len(bad)} bad cells of {diff.size}; first {bad[:5].tolist()} "
                      f"x-range {bad[:,0].min() if len(bad) else None}-{bad[:,0].max() if len(bad) else None} "
                      f"y {bad[:,1].min() if len(bad) else None}-{bad[:,1].max() if len(bad) else None} "
                      f"z {bad[:,2].min() if len(bad) else None}-{bad[:,2].max() if len(bad) else None}")
    del os.environ["INSB200_FUSED_STOP_AFTER"]
    sp = I.create_stepper(m_g, setup=sg, psolver=ps_g, state={"u": I.to_device(sg, u0)}, t=0.0)
    sp = I.timestep_(m_g, I.navierstokes_, sp, dt, params=dict(viscosity=nu))
    print("full step:", rel(I.to_host(sp["state"]["u"]), u))


if __name__ == "__main__":
    main()
