"""Multi-rank tests of the sharded paths (row-sharded gbmv, RHS-sharded solve, column-sharded banded x banded).  CPU: world_size-2 gloo run of the host-side plumbing (static data halo,
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


def test_gbmm_shard_geometry_covers_every_entry():
    """Every in-band entry of C[:, j0:j1] needs only A's columns [v0, v1) and rows [r0, r1); the relabelled bandwidths are
    non-negative and add up like the unsharded ones."""
    from bandedmatrices_b200.sharded import gbmm_shard_geometry, rhs_bounds, shard_bounds

    n, Ab, Bb = 5000, (7, 3), (2, 9)
    for world in (1, 2, 3, 8):
        cols = []
        for r in range(world):
            j0, j1 = shard_bounds(n, r, world)
            g = gbmm_shard_geometry(n, Ab, Bb, j0, j1)
            assert g["v0"] == max(0, j0 - Bb[1]) and g["v1"] == min(n, j1 + Bb[0])
            assert g["r0"] == max(0, j0 - Ab[1] - Bb[1]) and g["r1"] == min(n, j1 + Ab[0] + Bb[0])
            assert g["C"] == (g["A"][0] + g["B"][0], g["A"][1] + g["B"][1]) and min(g["A"] + g["B"] + g["C"]) >= 0
            cols += list(range(j0, j1))
        assert cols == list(range(n))
        blocks = [rhs_bounds(37, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == 37 and all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_gbmm_fake_ranks_one_gpu(world):
    """The column-sharded banded x banded product, every rank's slab computed in turn on ONE GPU: the slabs together are the
    unsharded bmb200_dgbmm_bb result bit for bit (tensor-core tile kernel at (32,32), sweep kernel at narrow bands)."""
    import numpy as np
    import torch

    import bandedmatrices_b200 as bm
    from bandedmatrices_b200.sharded import ShardedGbmm, shard_bounds

    for (n, Ab, Bb) in [(40000, (32, 32), (32, 32)), (9000, (4, 3), (4, 3)), (12000, (20, 9), (11, 16))]:
        A, B = bm.brand(n, n, *Ab, seed=5), bm.brand(n, n, *Bb, seed=6)
        Cl, Cu = Ab[0] + Bb[0], Ab[1] + Bb[1]
        Cm = bm.BandedMatrix.undef((n, n), (Cl, Cu))
        bm.mul_(Cm, A, B)
        for r in range(world):
            j0, j1 = shard_bounds(n, r, world)
            op = ShardedGbmm(n, Ab, Bb, j0, j1, A.data[max(0, j0 - Bb[1]): min(n, j1 + Bb[0])])
            C_local = torch.full((j1 - j0, Cl + Cu + 1), float("nan"), dtype=torch.float64, device="cuda")
            op(1.0, B.data[j0:j1], 0.0, C_local)
            jj = torch.arange(j0, j1, device="cuda").unsqueeze(1)
            rr = torch.arange(Cl + Cu + 1, device="cuda").unsqueeze(0)
            inm = (jj + rr - Cu >= 0) & (jj + rr - Cu < n)
            assert torch.equal(C_local[inm].view(torch.int64), Cm.data[j0:j1][inm].view(torch.int64)), (n, Ab, Bb, r)


@pytest.mark.gpu
def test_sharded_solve_fake_ranks_one_gpu(oracle_c, rng):
    """RHS-sharded ldiv!: the column blocks solved one after the other on one GPU equal DGBTRS on the whole B."""
    import numpy as np

    import bandedmatrices_b200 as bm
    import oracle
    from bandedmatrices_b200.sharded import ShardedSolve, rhs_bounds

    n, l, u, nrhs, world = 5000, 16, 16, 37, 4
    A = oracle.brand(rng, n, n, l, u)
    ab, ipiv, info = oracle.lu(oracle_c, A)
    B = np.asfortranarray(rng.standard_normal((n, nrhs)))
    ref = B.copy(order="F")
    oracle.ldiv(oracle_c, "N", ab, ipiv, l, u, ref)
    F = bm.lu(bm.BandedMatrix.from_banddata(A.data, n, l, u))
    for r in range(world):
        q0, q1 = rhs_bounds(nrhs, r, world)
        S = ShardedSolve(F, n, l, u, r, 1)  # world 1: no broadcast
        X = bm.to_colmajor(B[:, q0:q1])
        S.ldiv_(X)
        assert np.array_equal(X.cpu().numpy(), ref[:, q0:q1])


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
