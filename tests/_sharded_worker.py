"""Worker for the sharded-gbmv tests (launched with torch.distributed.run; one rank per process).
backend gloo  (CPU, here):   exercises shard_bounds / slab_geometry / build_extended_slab and checks, with the oracle
                             as the arithmetic, that the slab decomposition reproduces the global gbmv bit for bit.
backend nccl  (GPU box):     the real thing -- bmb200_dgbmv_sharded with the in-kernel NVLink halo push -- against
                             the oracle's global result, bit for bit, over several calls (epoch parity) and shapes."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from bandedmatrices_b200.sharded import (broadcast_factors, build_extended_slab, gbmm_shard_geometry, rhs_bounds,  # noqa: E402
                                         shard_bounds, slab_geometry)


def check_sharded_solve(backend, rank, world, C):
    """RHS-sharded ldiv!: factor on rank 0, broadcast AB + ipiv, every rank solves its block of columns;
    the blocks together must be the oracle's DGBTRS solution bit for bit."""
    n, l, u, nrhs = 3000, 5, 4, 7 * world + 3
    rng = np.random.default_rng(11)
    A = oracle.brand(rng, n, n, l, u)
    B = np.asfortranarray(rng.standard_normal((n, nrhs)))
    ab, ipiv, info = oracle.lu(C, A)
    ref = B.copy(order="F")
    oracle.ldiv(C, "N", ab, ipiv, l, u, ref)
    q0, q1 = rhs_bounds(nrhs, rank, world)
    owned = [rhs_bounds(nrhs, r, world) for r in range(world)]
    assert owned[0][0] == 0 and owned[-1][1] == nrhs and all(owned[i][1] == owned[i + 1][0] for i in range(world - 1))
    if backend == "gloo":
        # plumbing: only rank 0 holds the factors before the broadcast
        data = torch.as_tensor(np.ascontiguousarray(ab.T)) if rank == 0 else torch.zeros((n, 2 * l + u + 1), dtype=torch.float64)
        piv = torch.as_tensor(ipiv.astype(np.int64)) if rank == 0 else torch.zeros(n, dtype=torch.int64)
        broadcast_factors(data, piv, 0)
        assert np.array_equal(data.numpy().T, ab) and np.array_equal(piv.numpy(), ipiv)
        Xl = np.asfortranarray(B[:, q0:q1].copy())
        oracle.ldiv(C, "N", np.asfortranarray(data.numpy().T), piv.numpy(), l, u, Xl)
        assert np.array_equal(Xl, ref[:, q0:q1])
    else:
        import bandedmatrices_b200 as bm
        from bandedmatrices_b200.sharded import ShardedSolve

        F = bm.lu(bm.BandedMatrix.from_banddata(A.data, n, l, u)) if rank == 0 else None
        S = ShardedSolve(F, n, l, u, rank, world)
        assert np.array_equal(S.F.ipiv, ipiv)
        Xl = bm.to_colmajor(B[:, q0:q1])
        S.ldiv_(Xl)
        assert np.array_equal(Xl.cpu().numpy(), ref[:, q0:q1]), ("sharded solve", rank)


def check_sharded_cholesky_solve(backend, rank, world, C):
    """RHS-sharded ldiv! of the banded Cholesky: factor triangle broadcast from rank 0, every rank solves its block of columns;
    the blocks together are the oracle's DPBTRS solution (to rounding: the device runs both sweeps as column sweeps)."""
    for uplo, n, kd in (("U", 2500, 4), ("L", 1800, 70)):
        nrhs = 5 * world + 2
        rng = np.random.default_rng(17)
        ab = np.asfortranarray(rng.standard_normal((kd + 1, n)))
        ab[kd if uplo == "U" else 0, :] = 2.0 * (kd + 1) + rng.random(n)
        B = np.asfortranarray(rng.standard_normal((n, nrhs)))
        fac = ab.copy(order="F")
        assert C.pbtrf(uplo, n, kd, fac, kd + 1) == 0
        ref = B.copy(order="F")
        assert C.pbtrs(uplo, n, kd, nrhs, fac, kd + 1, ref, n) == 0
        q0, q1 = rhs_bounds(nrhs, rank, world)
        if backend == "gloo":
            tri = torch.as_tensor(np.ascontiguousarray(fac.T)) if rank == 0 else torch.zeros((n, kd + 1), dtype=torch.float64)
            dist.broadcast(tri, 0)
            assert np.array_equal(tri.numpy().T, fac)
            Xl = np.asfortranarray(B[:, q0:q1].copy())
            assert C.pbtrs(uplo, n, kd, q1 - q0, np.asfortranarray(tri.numpy().T), kd + 1, Xl, n) == 0
            assert np.array_equal(Xl, ref[:, q0:q1])
        else:
            import bandedmatrices_b200 as bm
            from bandedmatrices_b200.sharded import ShardedCholeskySolve

            tri = None
            if rank == 0:
                tri = torch.as_tensor(np.ascontiguousarray(ab.T)).cuda()
                _, info = bm.pbtrf_(uplo, n, kd, tri)
                assert info == 0
            S = ShardedCholeskySolve(tri, uplo, n, kd, rank, world)
            if kd <= 64:
                assert np.array_equal(S.tri.cpu().numpy().T, fac), "broadcast Cholesky factor differs from DPBTF2"
            Xl = bm.to_colmajor(B[:, q0:q1])
            S.ldiv_(Xl)
            assert np.max(np.abs(Xl.cpu().numpy() - ref[:, q0:q1])) <= 1e-12 * np.max(np.abs(ref)), ("sharded Cholesky solve", rank)


def check_sharded_gbmm(backend, rank, world, C):
    """Column-sharded banded x banded: every rank's slab of C equals the same columns of the unsharded _gbmm! result."""
    for (n, Ab, Bb) in [(1500, (3, 2), (4, 1)), (2000, (9, 12), (10, 9)), (4096, (32, 32), (32, 32))]:
        rng = np.random.default_rng(13)
        A = oracle.brand(rng, n, n, *Ab)
        B = oracle.brand(rng, n, n, *Bb)
        Cl, Cu = Ab[0] + Bb[0], Ab[1] + Bb[1]
        ref = np.zeros((Cl + Cu + 1, n), order="F")
        oracle.gbmm_kernel(C, 1.0, A.data, B.data, 0.0, ref, n, n, n, Ab[0], Ab[1], Bb[0], Bb[1], Cl, Cu)
        j0, j1 = shard_bounds(n, rank, world)
        g = gbmm_shard_geometry(n, Ab, Bb, j0, j1)
        assert g["A"][0] + g["A"][1] == sum(Ab) and g["B"][0] + g["B"][1] == sum(Bb) and g["C"][0] + g["C"][1] == Cl + Cu
        assert min(g["A"] + g["B"] + g["C"]) >= 0
        dev = "cuda" if backend == "nccl" else "cpu"
        A_local = torch.as_tensor(np.ascontiguousarray(A.data.T[j0:j1])).to(dev)
        A_cols = build_extended_slab(A_local, n, j0, j1, Bb[1], Bb[0], rank, world)
        assert np.array_equal(A_cols.cpu().numpy(), A.data.T[g["v0"]:g["v1"]])
        # in-matrix entries of the slab's columns
        rr, jj = np.meshgrid(np.arange(Cl + Cu + 1), np.arange(j0, j1), indexing="ij")
        inm = (jj + rr - Cu >= 0) & (jj + rr - Cu < n)
        if backend == "gloo":
            sub = np.zeros((Cl + Cu + 1, j1 - j0), order="F")
            oracle.gbmm_kernel(C, 1.0, np.asfortranarray(A_cols.numpy().T), np.asfortranarray(B.data[:, j0:j1]), 0.0, sub, g["rows"],
                               g["inner"], g["cols"], g["A"][0], g["A"][1], g["B"][0], g["B"][1], g["C"][0], g["C"][1])
            assert np.array_equal(sub[inm], ref[:, j0:j1][inm]), ("gbmm slab", n, rank)
        else:
            from bandedmatrices_b200.sharded import ShardedGbmm

            op = ShardedGbmm(n, Ab, Bb, j0, j1, A_local, rank, world, extend=True)
            B_local = torch.as_tensor(np.ascontiguousarray(B.data.T[j0:j1])).cuda()
            C_local = torch.full((j1 - j0, Cl + Cu + 1), float("nan"), dtype=torch.float64, device="cuda")
            op(1.0, B_local, 0.0, C_local)
            torch.cuda.synchronize()
            assert np.array_equal(C_local.cpu().numpy().T[inm], ref[:, j0:j1][inm]), ("gbmm slab", n, rank)


def main():
    backend = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    else:
        dist.init_process_group("gloo")
    C = oracle.backend("C")
    for (n, kl, ku) in [(4099, 4, 3), (1000, 1, 1), (20000, 7, 0), (5000, 0, 5), (3001, 2, 2), (70000, 4, 3)]:
        rng = np.random.default_rng(7)  # same global problem on every rank
        A = oracle.brand(rng, n, n, kl, ku, corners=np.nan)
        lda = kl + ku + 1
        c0, c1 = shard_bounds(n, rank, world)
        geo = slab_geometry(n, c0, c1, kl, ku)
        dev = "cuda" if backend == "nccl" else "cpu"
        local = torch.as_tensor(np.ascontiguousarray(A.data.T[c0:c1])).to(dev)
        ext = build_extended_slab(local, n, c0, c1, kl, ku, rank, world)
        assert ext.shape == (geo["ns"], lda)
        want = A.data.T[geo["cs"]:geo["ce"]]
        assert np.array_equal(ext.cpu().numpy(), want, equal_nan=True), "static data halo wrong"
        for it in range(3):  # several calls: exercises the epoch / parity double buffering
            x = rng.standard_normal(n)
            y0 = rng.standard_normal(n)
            alpha, beta = (1.0, 0.0) if it == 0 else (0.5 + it, 0.25 * it)
            ref = y0.copy()
            oracle.gbmv(C, "N", n, kl, ku, alpha, A.data, x, beta, ref)
            if backend == "gloo":
                # slab arithmetic with the oracle on the extended slab and the x window [cs, ce)
                yl = y0[c0:c1].copy()
                sub = np.asfortranarray(ext.numpy().T)
                oracle.gbmv(C, "N", geo["ms"], geo["kls"], geo["kus"], alpha, sub, x[geo["cs"]:geo["ce"]].copy(), beta, yl)
                assert np.array_equal(yl, ref[c0:c1]), (n, kl, ku, rank)
            else:
                import bandedmatrices_b200 as bm
                from bandedmatrices_b200.sharded import ShardedGbmv

                if it == 0:
                    op = ShardedGbmv(n, c0, c1, kl, ku, bm.BandedMatrix(local, c1 - c0, kl, ku), rank, world)
                xl = torch.as_tensor(x[c0:c1].copy()).cuda()
                yl = torch.as_tensor(y0[c0:c1].copy()).cuda()
                op(alpha, xl, beta, yl)
                op.hd.sync()
                assert np.array_equal(yl.cpu().numpy(), ref[c0:c1]), (n, kl, ku, rank, it)
        if backend == "nccl":
            op.close()
    check_sharded_solve(backend, rank, world, C)
    check_sharded_cholesky_solve(backend, rank, world, C)
    check_sharded_gbmm(backend, rank, world, C)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("sharded worker ok", backend, world)


if __name__ == "__main__":
    main()
