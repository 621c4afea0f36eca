"""P-GPU == 1-GPU equality -- launches tests/mgpu_worker.py under torchrun on
all visible GPUs (needs >= 2)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("grid,dtype,case", [((64, 32, 64), "f64", "periodic"), ((128, 64, 128), "f32", "periodic"),
                                             ((64, 24, 64), "f64", "channel"), ((128, 32, 64), "f32", "channel"),
                                             ((64, 32, 64), "f64", "rb"), ((128, 64, 128), "f32", "rb"),
                                             ((64, 32, 64), "f64", "all")])
def test_slab_equals_single_gpu(grid, dtype, case):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 1 << (n.bit_length() - 1)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "mgpu_worker.py"),
           "--grid", *map(str, grid), "--dtype", dtype, "--case", case, "--strict"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines and all(l["ok"] for l in lines), r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
def test_missing_peer_is_an_error_not_a_hang():
    """A rank that never joins an exchange makes its neighbours' waiting kernels trap after the time-out
    (INSB200_PEER_TIMEOUT_MS): a loud, sticky error instead of a hang or stale ghost planes."""
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29534", os.path.join(ROOT, "tests", "mgpu_worker.py"), "--case", "timeout"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, INSB200_PEER_TIMEOUT_MS="1500"))
    assert r.returncode == 0 and '"peer_timeout"' in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
