"""Helpers for the GPU parity tests: build the same problem for the oracle (NumPy) and for
the CUDA path (insb200 over the C ABI) and move fields between them."""
import numpy as np

import oracle as O


def _bc_pair(kind, mod):
    if isinstance(kind, tuple):
        return tuple(_bc_one(k, mod) for k in kind)
    return (_bc_one(kind, mod), _bc_one(kind, mod))


def _bc_one(k, mod):
    if isinstance(k, tuple):  # ("D", value)
        return mod.DirichletBC(k[1])
    return {"P": mod.PeriodicBC, "D": mod.DirichletBC, "S": mod.SymmetricBC, "O": mod.PressureBC}[k]()


def make_setups(x, bcs, temp_bcs=None, T=np.float64):
    """bcs: per-dim spec, 'P' | 'D' | 'S' | 'O' | (left, right) | ('D', value)."""
    import insb200

    xs = tuple(np.asarray(v).astype(T) for v in x)
    out = []
    for mod in (O, insb200):
        d = {"u": tuple(_bc_pair(k, mod) for k in bcs)}
        if temp_bcs is not None:
            d["temp"] = tuple(_bc_pair(k, mod) for k in temp_bcs)
        out.append(mod.Setup(xs, d))
    return out[0], out[1]


def uniform(n, L=1.0):
    return np.linspace(0.0, L, n + 1)


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def random_vec(setup_o, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    return np.asfortranarray((scale * rng.standard_normal(setup_o.N + (setup_o.D,))).astype(setup_o.T))


def random_scal(setup_o, seed):
    rng = np.random.default_rng(seed)
    return np.asfortranarray(rng.standard_normal(setup_o.N).astype(setup_o.T))


def smooth_vec(setup_o, seed=0):
    """Smooth O(1) field (sum of a few modes) sampled at the velocity points + noise 1%."""
    from oracle import bc

    rng = np.random.default_rng(seed)
    D = setup_o.D
    u = O.vectorfield(setup_o)
    for a in range(D):
        acc = 0
        for m in range(3):
            ph = rng.uniform(0, 2 * np.pi, D)
            kk = rng.integers(1, 4, D)
            term = 1.0
            for b in range(D):
                shape = [1] * D
                shape[b] = -1
                L = float(setup_o.xlims[b][1] - setup_o.xlims[b][0])
                term = term * np.sin(2 * np.pi * kk[b] * setup_o.xu[a][b].reshape(shape) / L + ph[b])
            acc = acc + term / (m + 1)
        u[..., a] = acc.astype(setup_o.T)
    u += (0.01 * rng.standard_normal(u.shape)).astype(setup_o.T)
    bc.apply_bc_u(u, 0.0, setup_o)
    return u
