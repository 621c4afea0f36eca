"""Host-side logic of the z-slab decomposition, run on CPU with world_size 2 over gloo:
unique-id sharing, slab index maps, scatter -> gather round trip.  (The GPU equality test
P-GPU == 1-GPU is tests/mgpu_worker.py, driven by test_multi_gpu.py.)"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "incompressiblenavierstokes.jl_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import insb200 as I
    from insb200 import api

    try:
        nid = api.share_unique_id()
        ids = [None] * world
        dist.all_gather_object(ids, nid)
        assert len(nid) == 128 and all(i == nid for i in ids)
        n = (8, 6, 12)
        x = tuple(np.linspace(0.0, 1.0, m + 1) for m in n)
        bcs = {"u": ((I.PeriodicBC(), I.PeriodicBC()),) * 3}
        s = I.Setup(x, bcs, rank=rank, world=world, nccl_id=nid)
        assert s.N == (10, 8, 14) and s.N_local == (10, 8, 12 // world + 2) and s.zoff == rank * 12 // world
        # ownership: the interior z planes of all ranks tile the global interior exactly once
        own = np.zeros(s.N[2], int)
        lo, hi = s.local_range(s.Ip[2], 2)
        mine = np.zeros(s.N[2], int)
        mine[lo + s.zoff:hi + s.zoff] = 1
        t = torch.from_numpy(mine)
        dist.all_reduce(t)
        own = t.numpy()
        assert own[0] == 0 and own[-1] == 0 and np.all(own[1:-1] == 1)
        # scatter -> gather round trip on a ghost-consistent (periodic) global array
        rng = np.random.default_rng(3)
        g = rng.standard_normal(s.N + (3,))
        g[:, :, 0] = g[:, :, -2]
        g[:, :, -1] = g[:, :, 1]
        loc = api.local_slab(s, g)
        assert loc.shape == s.N_local + (3,)
        tl = torch.from_numpy(np.ascontiguousarray(np.asfortranarray(loc).T))
        back = api.gather_global(s, tl)
        assert np.array_equal(back, g)
        # velocity sampling points of this slab are the global ones shifted by zoff
        lo, hi = s.local_range(s.Iu[2][2], 2)
        assert np.array_equal(s.xu[2][2][lo + s.zoff:hi + s.zoff], x[2][1:][s.zoff:s.zoff + s.nzl])
        # bad decompositions are refused
        with pytest.raises(ValueError):
            I.Setup((x[0], x[1], np.linspace(0, 1, 8)), bcs, rank=rank, world=world, nccl_id=nid)
        with pytest.raises(ValueError):  # a pressure boundary in the decomposed direction shifts the DOF ranges
            I.Setup(x, {"u": ((I.PeriodicBC(), I.PeriodicBC()),) * 2 + ((I.DirichletBC(), I.PressureBC()),)},
                    rank=rank, world=world, nccl_id=nid)
        with pytest.raises(ValueError):
            I.Setup(x, {"u": ((I.PeriodicBC(), I.PeriodicBC()),) * 2 + ((I.PeriodicBC(), I.DirichletBC()),)},
                    rank=rank, world=world, nccl_id=nid)
        # walls in the decomposed direction (BASELINE config 4): the wall-normal velocity DOFs exclude the last
        # interior plane, which only the last rank owns; the other DOF ranges tile the interior as before
        sw = I.Setup(x, {"u": ((I.PeriodicBC(), I.PeriodicBC()),) * 2 + ((I.DirichletBC(), I.DirichletBC()),),
                         "temp": ((I.PeriodicBC(), I.PeriodicBC()),) * 2 + ((I.DirichletBC(1.0), I.DirichletBC(0.0)),)},
                     rank=rank, world=world, nccl_id=nid)
        lo, hi = sw.local_range(sw.Iu[2][2], 2)
        assert (lo, hi) == (1, sw.nzl + 1 - (1 if rank == world - 1 else 0))
        assert sw.local_range(sw.Iu[0][2], 2) == (1, sw.nzl + 1) and sw.local_range(sw.Ip[2], 2) == (1, sw.nzl + 1)
        q.put((rank, "ok"))
    except BaseException as e:  # noqa: BLE001  (pytest's Failed derives from BaseException)
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_slab_host_logic_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
