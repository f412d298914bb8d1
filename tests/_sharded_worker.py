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
from bandedmatrices_b200.sharded import build_extended_slab, shard_bounds, slab_geometry  # noqa: E402


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
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("sharded worker ok", backend, world)


if __name__ == "__main__":
    main()
