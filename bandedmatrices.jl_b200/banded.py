"""Device-resident ``BandedMatrix`` -- the data model of the hot path.

Mirrors ``BandedMatrix{Float64,<device array>}`` of the reference:
struct src/banded/BandedMatrix.jl:16-28, ``_BandedMatrix`` check :21-27, ``brand`` :184-187,
``bandeddata``/``bandwidths`` :408-409, storage rule ``data[u+k-j+1, j] = A[k,j]`` :414-419,
sub-views ``bandeddata(view)`` :947-951.

Storage.  The reference keeps a column-major ``(l+u+1) x n`` array.  Here ``data`` is a CUDA
``torch.float64`` tensor of shape ``(n, l+u+1)`` whose memory is exactly that array
(``data[j, r]`` is band row ``r`` of column ``j``; ``data.stride() == (lda, 1)``), so
``data.data_ptr()`` / ``lda`` go straight into the C ABI and torch slicing gives the
pointer arithmetic for sub-views.  torch is only the allocator here.
"""
from __future__ import annotations

import numpy as np
import torch


class DimensionMismatch(ValueError):
    """Julia's ``DimensionMismatch``."""


class BandError(Exception):
    """src/generic/Band.jl:75-89 -- thrown when C has too few bands for A*B (gbmm.jl:252,267)."""

    def __init__(self, A, band):
        super().__init__(f"attempt to access {A.m}x{A.n} BandedMatrix with bandwidths {(A.l, A.u)} at band {band}")
        self.band = band


class LAPACKException(Exception):
    """``LAPACK.chklapackerror(info)`` for info > 0 (zero pivot in gbtrf!)."""

    def __init__(self, info):
        super().__init__(f"LAPACKException({info})")
        self.info = info


def _device(device=None) -> torch.device:
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        raise RuntimeError("bandedmatrices.jl_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class BandedMatrix:
    """``_BandedMatrix(data, m, l, u)``: m x n banded matrix with bandwidths (l, u)."""

    __slots__ = ("data", "m", "l", "u")

    def __init__(self, data: torch.Tensor, m: int, l: int, u: int):
        if data.dim() != 2 or data.dtype != torch.float64:
            raise TypeError("data must be a 2-D float64 tensor of shape (n, l+u+1)")
        if data.shape[1] != max(0, l + u + 1):  # BandedMatrix.jl:22-24
            raise ValueError("Data matrix must have number rows equal to number of bands")
        if data.shape[1] > 1 and data.stride(1) != 1:
            raise ValueError("band rows of one column must be contiguous")
        self.data, self.m, self.l, self.u = data, int(m), int(l), int(u)

    # ---- constructors -------------------------------------------------------------------------
    @classmethod
    def undef(cls, shape, bandwidths, device=None) -> "BandedMatrix":
        """BandedMatrix{Float64}(undef, (m,n), (l,u)) -- BandedMatrix.jl:51-68."""
        (m, n), (l, u) = shape, bandwidths
        return cls(torch.empty((n, max(0, l + u + 1)), dtype=torch.float64, device=_device(device)), m, l, u)

    @classmethod
    def zeros(cls, shape, bandwidths, device=None) -> "BandedMatrix":
        (m, n), (l, u) = shape, bandwidths
        return cls(torch.zeros((n, max(0, l + u + 1)), dtype=torch.float64, device=_device(device)), m, l, u)

    @classmethod
    def from_banddata(cls, data_cm, m: int, l: int, u: int, device=None) -> "BandedMatrix":
        """From a host/array ``(l+u+1) x n`` band-storage array (the reference's ``A.data``)."""
        t = torch.as_tensor(np.ascontiguousarray(np.asarray(data_cm, dtype=np.float64).T))
        return cls(t.to(_device(device)), m, l, u)

    @classmethod
    def from_dense(cls, A, bandwidths, device=None) -> "BandedMatrix":
        """BandedMatrix(A, (l,u)) -- BandedMatrix.jl:222-232 (entries outside the band are dropped)."""
        A = np.asarray(A, dtype=np.float64)
        m, n = A.shape
        l, u = bandwidths
        data = np.zeros((max(0, l + u + 1), n))
        for r in range(data.shape[0]):
            off = u - r  # band row r holds diagonal k - j = -off ... k = j - off
            js = np.arange(max(0, off), min(n, m + off))
            if js.size:
                data[r, js] = A[js - off, js]
        return cls.from_banddata(data, m, l, u, device)

    # ---- queries ------------------------------------------------------------------------------
    @property
    def n(self) -> int:
        return int(self.data.shape[0])

    @property
    def shape(self):
        return (self.m, self.n)

    @property
    def lda(self) -> int:
        return int(self.data.stride(0)) if self.data.shape[0] > 1 else max(1, int(self.data.shape[1]))

    @property
    def ptr(self) -> int:
        return self.data.data_ptr()

    @property
    def T(self) -> "Transposed":
        return Transposed(self)

    def banddata_host(self) -> np.ndarray:
        """``Array(bandeddata(A))`` on the host, shape (l+u+1, n), Fortran order."""
        return np.asfortranarray(self.data.detach().cpu().numpy().T)

    def to_dense(self) -> np.ndarray:
        """``Matrix(A)`` on the host (tests / debugging only)."""
        d = self.banddata_host()
        out = np.zeros((self.m, self.n))
        for r in range(d.shape[0]):
            off = self.u - r
            js = np.arange(max(0, off), min(self.n, self.m + off))
            if js.size:
                out[js - off, js] = d[r, js]
        return out

    def copy(self) -> "BandedMatrix":
        return BandedMatrix(self.data.clone(), self.m, self.l, self.u)

    # ---- views (BandedSubBandedMatrix, BandedMatrix.jl:918-951): same band rows, shifted bandwidths ----
    def view_cols(self, c0: int, c1: int | None = None) -> "BandedMatrix":
        """view(A, :, c0+1:c1): V[k,j'] = A[k, j'+c0]  ->  bandwidths (l+c0, u-c0)."""
        c1 = self.n if c1 is None else c1
        return BandedMatrix(self.data[c0:c1], self.m, self.l + c0, self.u - c0)

    def view_rows(self, r0: int, r1: int | None = None) -> "BandedMatrix":
        """view(A, r0+1:r1, :): V[k',j] = A[k'+r0, j]  ->  bandwidths (l-r0, u+r0)."""
        r1 = self.m if r1 is None else r1
        return BandedMatrix(self.data, r1 - r0, self.l - r0, self.u + r0)

    def __repr__(self):
        return f"BandedMatrix({self.m}x{self.n}, bandwidths=({self.l},{self.u}), device={self.data.device})"


class Transposed:
    """``transpose(A)`` / ``A'`` of a real BandedMatrix (BandedRowMajor layout, matmul.jl:82-86)."""

    __slots__ = ("parent",)

    def __init__(self, parent: BandedMatrix):
        self.parent = parent

    @property
    def shape(self):
        return (self.parent.n, self.parent.m)

    @property
    def T(self):
        return self.parent


def bandwidths(A):
    if isinstance(A, Transposed):
        return (A.parent.u, A.parent.l)
    return (A.l, A.u)


def bandwidth(A, i: int):
    return bandwidths(A)[i - 1]


def bandeddata(A: BandedMatrix) -> torch.Tensor:
    """Column-major ``(l+u+1) x n`` view of the band storage (BandedMatrix.jl:408)."""
    return A.data.T


def brand(m: int, n: int, l: int, u: int, seed: int | None = None, device=None) -> BandedMatrix:
    """brand(m,n,l,u): uniform [0,1) over the WHOLE data array, corner slots included (BandedMatrix.jl:184-187)."""
    dev = _device(device)
    g = torch.Generator(device=dev)
    if seed is not None:
        g.manual_seed(int(seed))
    else:
        g.seed()  # a fresh Generator starts from torch's fixed default seed: draw a new one, like the reference's global RNG
    data = torch.rand((n, max(0, l + u + 1)), dtype=torch.float64, device=dev, generator=g)
    return BandedMatrix(data, m, l, u)


def colmajor(rows: int, cols: int, device=None, fill: float | None = None) -> torch.Tensor:
    """A Julia-style column-major dense ``rows x cols`` matrix (stride (1, rows)) on the device."""
    base = torch.empty((cols, rows), dtype=torch.float64, device=_device(device))
    if fill is not None:
        base.fill_(fill)
    return base.T


def to_colmajor(X, device=None) -> torch.Tensor:
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        return torch.as_tensor(X.copy()).to(_device(device))
    return torch.as_tensor(np.ascontiguousarray(X.T)).to(_device(device)).T
