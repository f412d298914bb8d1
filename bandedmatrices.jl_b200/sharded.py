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
