"""Row-sharded multi-GPU gbmv: one process per GPU, torch.distributed for the plumbing.

Rank r owns rows [c0, c1) of the square n x n band matrix.  The (kl+ku)-column DATA halo is static and is
replicated once here (an all_gather of a few columns); the x halo moves inside the CUDA kernel itself through
NVLink peer stores into mailboxes opened with CUDA IPC (csrc/sharded.cu) -- there is no collective on the data
path.  The reference has no distributed code; this is the B200-native scaling of its one-dgbmv_ matvec
(src/generic/matmul.jl:21-23)."""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib
from .banded import BandedMatrix

vp = C.c_void_p


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous row/column slab [c0, c1) of rank ``rank``."""
    return (n * rank) // world, (n * (rank + 1)) // world


def slab_geometry(n: int, c0: int, c1: int, kl: int, ku: int):
    """Columns [cs, ce) a slab needs and the bandwidths of A[c0:c1, cs:ce] in the same band storage."""
    cs, ce = max(0, c0 - kl), min(n, c1 + ku)
    hl, hr = c0 - cs, ce - c1
    return {"cs": cs, "ce": ce, "hl": hl, "hr": hr, "kls": kl - hl, "kus": ku + hl, "ms": c1 - c0, "ns": ce - cs}


def build_extended_slab(data_local: torch.Tensor, n: int, c0: int, c1: int, kl: int, ku: int, rank: int, world: int,
                        group=None) -> torch.Tensor:
    """Replicate the static data halo: returns band data for columns [cs, ce) (tensor layout (ncols, lda)).
    Works on any backend (NCCL on GPUs, gloo on CPU tensors in the unit tests)."""
    geo = slab_geometry(n, c0, c1, kl, ku)
    nl, lda = data_local.shape
    if world > 1 and nl < max(kl, ku):
        raise ValueError("every slab must hold at least max(kl,ku) columns")
    h = max(kl, ku, 1)
    edge = torch.zeros((2, h, lda), dtype=data_local.dtype, device=data_local.device)
    edge[0, : min(h, nl)] = data_local[: min(h, nl)]          # my first columns (someone's right halo)
    edge[1, h - min(h, nl) :] = data_local[nl - min(h, nl) :]  # my last columns  (someone's left halo)
    if world > 1:
        allg = [torch.empty_like(edge) for _ in range(world)]
        dist.all_gather(allg, edge, group=group)
    else:
        allg = [edge]
    parts = []
    if geo["hl"]:
        parts.append(allg[rank - 1][1, h - geo["hl"] :])
    parts.append(data_local)
    if geo["hr"]:
        parts.append(allg[rank + 1][0, : geo["hr"]])
    return torch.cat(parts, dim=0).contiguous() if len(parts) > 1 else data_local


class ShardedGbmv:
    """y_local <- alpha * A[c0:c1, :] * x + beta * y_local with x, y sharded like the rows."""

    def __init__(self, n: int, c0: int, c1: int, kl: int, ku: int, A_local: BandedMatrix, rank: int, world: int,
                 group=None, extended: bool = False):
        self.n, self.c0, self.c1, self.kl, self.ku, self.rank, self.world = n, c0, c1, kl, ku, rank, world
        dev = A_local.data.device
        self.hd = _lib.handle(dev.index)
        self.data_ext = A_local.data if extended else build_extended_slab(A_local.data, n, c0, c1, kl, ku, rank, world, group)
        self.lda = int(self.data_ext.stride(0))
        # mailboxes: create, exchange the CUDA IPC handles, open the neighbours'
        buf = (C.c_ubyte * 64)()
        self.hd.check(self.hd.lib.bmb200_halo_create(self.hd.h, max(kl, ku, 1), buf), "halo_create")
        mine = bytes(buf)
        handles = [None] * world
        if world > 1:
            dist.all_gather_object(handles, mine, group=group)
        else:
            handles = [mine]
        left = (C.c_ubyte * 64).from_buffer_copy(handles[rank - 1]) if rank > 0 else None
        right = (C.c_ubyte * 64).from_buffer_copy(handles[rank + 1]) if rank + 1 < world else None
        self.hd.check(self.hd.lib.bmb200_halo_connect(self.hd.h, rank, world, left, right), "halo_connect")
        if world > 1:
            dist.barrier(group=group)  # every mailbox is open before the first push

    def __call__(self, alpha: float, x: torch.Tensor, beta: float, y: torch.Tensor) -> torch.Tensor:
        hd = self.hd
        hd.set_stream(torch.cuda.current_stream(x.device).cuda_stream)
        hd.check(hd.lib.bmb200_dgbmv_sharded(hd.h, self.n, self.c0, self.c1, self.kl, self.ku, float(alpha),
                                             vp(self.data_ext.data_ptr()), self.lda, vp(x.data_ptr()), float(beta),
                                             vp(y.data_ptr())), "dgbmv_sharded")
        return y

    def close(self):
        self.hd.lib.bmb200_halo_destroy(self.hd.h)

    def check(self) -> None:
        """Synchronise and raise if a halo wait timed out (the kernel flags it and carries on with stale data)."""
        self.hd.sync()


# ---------------------------------------------------------------------------------------------------------------------
# LU / solve: the factorisation does not partition (ipiv must be bit-identical), the solve shards over RIGHT-HAND SIDES
# (reference site ldiv!(::BandedLU, B), src/banded/linalg.jl:24-30).  Factor once, broadcast the factors
# and the pivots, then every rank runs bmb200_dgbtrs on its own block of B's columns: no exchange during the solve.
# ---------------------------------------------------------------------------------------------------------------------
def rhs_bounds(nrhs: int, rank: int, world: int):
    """Contiguous block [q0, q1) of right-hand-side columns owned by ``rank``."""
    return (nrhs * rank) // world, (nrhs * (rank + 1)) // world


def broadcast_factors(data: torch.Tensor, ipiv: torch.Tensor, src: int = 0, group=None):
    """One broadcast of AB (the (n, 2l+u+1) factor slab) and one of ipiv from ``src`` (NCCL on GPUs, gloo on CPU
    tensors in the unit tests).  In place; returns the two tensors."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(data, src, group=group)
        dist.broadcast(ipiv, src, group=group)
    return data, ipiv


class ShardedSolve:
    """``ldiv!(F, B)`` with the columns of B sharded over the ranks.

    ``F`` is the BandedLU on rank ``src`` (``None`` elsewhere); ``n, l, u`` describe the ORIGINAL matrix (factors have
    bandwidths (l, l+u)).  After construction every rank holds a replica of the factors; ``ldiv_(B_local)`` solves
    this rank's column block in place."""

    def __init__(self, F, n: int, l: int, u: int, rank: int, world: int, group=None, src: int = 0, device=None):
        from .linalg import BandedLU

        self.rank, self.world, self.n, self.l, self.u = rank, world, n, l, u
        if F is not None:
            data, ipiv_d = F.factors.data, F.d_ipiv()
        else:
            dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
            data = torch.empty((n, 2 * l + u + 1), dtype=torch.float64, device=dev)
            ipiv_d = torch.empty(n, dtype=torch.int64, device=dev)
        broadcast_factors(data, ipiv_d, src, group)
        if F is None:
            F = BandedLU(BandedMatrix(data, n, l, l + u), ipiv_d.cpu().numpy(), 0, ipiv_d)
        self.F = F

    def ldiv_(self, B_local: torch.Tensor) -> torch.Tensor:
        from .linalg import ldiv_

        return ldiv_(self.F, B_local)


class ShardedCholeskySolve:
    """``ldiv!(cholesky(Symmetric(A)), B)`` (src/symbanded/BandedCholesky.jl:72-80) with the columns of B sharded over the ranks,
    like ``ShardedSolve``: the factorisation is one dependency chain and stays on rank ``src``; ONE broadcast of the
    (n, kd+1) factor triangle, then every rank runs ``bmb200_dpbtrs`` on its block of right-hand sides, no exchange."""

    def __init__(self, tri: torch.Tensor | None, uplo: str, n: int, kd: int, rank: int, world: int, group=None, src: int = 0, device=None):
        self.rank, self.world, self.n, self.kd, self.uplo = rank, world, n, kd, uplo
        if tri is None:
            dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
            tri = torch.empty((n, kd + 1), dtype=torch.float64, device=dev)
        elif not tri.is_contiguous():
            tri = tri.contiguous()
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.broadcast(tri, src, group=group)
        self.tri = tri

    def ldiv_(self, B_local: torch.Tensor) -> torch.Tensor:
        from .linalg import pbtrs_

        return pbtrs_(self.uplo, self.n, self.kd, self.tri, B_local)


# ---------------------------------------------------------------------------------------------------------------------
# banded x banded, sharded over the COLUMNS of B and C (reference _gbmm!, src/banded/gbmm.jl:296-340: column
# j of C is one gbmv over A's columns [j-Bu, j+Bl]).  Rank r owns columns [j0, j1) of B and C and needs A's columns
# [j0-Bu, j1+Bl): a STATIC halo, replicated once at distribution time (build_extended_slab) -- no exchange per product.
# In band storage an off-diagonal sub-block is the same data with re-labelled bandwidths, so the local product is one
# ordinary bmb200_dgbmm_bb call and every entry sees the same FMAs in the same order as in the unsharded product.
# ---------------------------------------------------------------------------------------------------------------------
def gbmm_shard_geometry(n: int, Ab, Bb, j0: int, j1: int):
    """Sub-problem of C[:, j0:j1] = A * B[:, j0:j1] for square n x n operands with bandwidths Ab=(Al,Au), Bb=(Bl,Bu) and
    C = (Al+Bl, Au+Bu): the A columns [v0, v1) and rows [r0, r1) it touches and the bandwidths of the three sub-blocks."""
    (Al, Au), (Bl, Bu) = Ab, Bb
    Cl, Cu = Al + Bl, Au + Bu
    v0, v1 = max(0, j0 - Bu), min(n, j1 + Bl)
    r0, r1 = max(0, j0 - Cu), min(n, j1 + Cl)
    sa, sb, sc = r0 - v0, v0 - j0, r0 - j0  # row-minus-column origin shift of each sub-block
    return {"v0": v0, "v1": v1, "r0": r0, "r1": r1, "rows": r1 - r0, "inner": v1 - v0, "cols": j1 - j0,
            "A": (Al - sa, Au + sa), "B": (Bl - sb, Bu + sb), "C": (Cl - sc, Cu + sc)}


class ShardedGbmm:
    """``mul!(C, A, B, alpha, beta)`` for banded A, B, C with B and C sharded by columns.

    ``A_cols`` holds A's data columns [v0, v1) (tensor layout (v1-v0, Al+Au+1)) -- pass ``A_local`` = columns [j0, j1)
    plus ``extend=True`` to let the constructor fetch the halo from the neighbours."""

    def __init__(self, n: int, Ab, Bb, j0: int, j1: int, A_cols: torch.Tensor, rank: int = 0, world: int = 1, group=None,
                 extend: bool = False):
        self.n, self.Ab, self.Bb, self.j0, self.j1 = n, tuple(Ab), tuple(Bb), j0, j1
        self.geo = gbmm_shard_geometry(n, Ab, Bb, j0, j1)
        if extend:
            A_cols = build_extended_slab(A_cols, n, j0, j1, Bb[1], Bb[0], rank, world, group)  # halo: Bu left, Bl right
        assert A_cols.shape[0] == self.geo["inner"], "A_cols must hold A's columns [j0-Bu, j1+Bl)"
        self.A_cols = A_cols
        g = self.geo
        if min(g["rows"], g["cols"]) - 1 < max(g["C"]):
            raise ValueError("slab narrower than the product's bandwidth: use fewer ranks")

    def __call__(self, alpha: float, B_local: torch.Tensor, beta: float, C_local: torch.Tensor) -> torch.Tensor:
        """B_local: (j1-j0, Bl+Bu+1) band data of B's columns; C_local: (j1-j0, Cl+Cu+1), overwritten."""
        g = self.geo
        hd = _lib.handle(C_local.device.index)
        lda = int(self.A_cols.stride(0))
        hd.check(hd.lib.bmb200_dgbmm_bb(hd.h, g["rows"], g["inner"], g["cols"], g["A"][0], g["A"][1], g["B"][0], g["B"][1],
                                        g["C"][0], g["C"][1], float(alpha), vp(self.A_cols.data_ptr()), lda,
                                        vp(B_local.data_ptr()), int(B_local.stride(0)), float(beta),
                                        vp(C_local.data_ptr()), int(C_local.stride(0))), "dgbmm_bb (column shard)")
        return C_local


class ShardedGbmmDense:
    """banded x dense with the RIGHT-HAND-SIDE columns of B and C sharded (src/generic/matmul.jl:243-256: one gbmv per
    column, so column blocks are independent); A is replicated, no halo and no exchange."""

    def __init__(self, A: BandedMatrix, rank: int = 0, world: int = 1):
        self.A, self.rank, self.world = A, rank, world

    def __call__(self, alpha: float, B_local: torch.Tensor, beta: float, C_local: torch.Tensor) -> torch.Tensor:
        from .linalg import mul_

        return mul_(C_local, self.A, B_local, alpha, beta)
