"""Multi-rank tests of the row-sharded gbmv.  CPU: world_size-2 gloo run of the host-side plumbing (static data halo,
slab geometry) with the oracle as arithmetic.  GPU: world_size-2 NCCL run of bmb200_dgbmv_sharded (needs >= 2 GPUs;
skipped on a 1-GPU box -- run it with `gpurun --gpus 2`)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(backend, nproc):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "_sharded_worker.py"), backend]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "sharded worker ok" in r.stdout


def test_sharded_plumbing_gloo_world2():
    _run("gloo", 2)


def test_slab_geometry():
    from bandedmatrices_b200.sharded import shard_bounds, slab_geometry

    n, kl, ku = 1000, 4, 3
    covered = []
    for r in range(8):
        c0, c1 = shard_bounds(n, r, 8)
        g = slab_geometry(n, c0, c1, kl, ku)
        assert g["kls"] + g["kus"] == kl + ku and g["kls"] >= 0 and g["kus"] >= 0
        assert g["hl"] == (kl if r > 0 else 0) and g["hr"] == (ku if r < 7 else 0)
        covered += list(range(c0, c1))
    assert covered == list(range(n))


@pytest.mark.gpu
def test_sharded_gbmv_nccl_world2():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    _run("nccl", 2)


@pytest.mark.gpu
def test_sharded_gbmv_nccl_world4():
    """Four ranks: interior ranks have two DIFFERENT neighbours (at world size 2 both mailboxes belong to the same peer)."""
    import torch

    if torch.cuda.device_count() < 4:
        pytest.skip("needs >= 4 GPUs (gpurun --gpus 4)")
    _run("nccl", 4)
