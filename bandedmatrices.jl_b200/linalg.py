"""Host-side mirror of the reference's hot-path drivers (``mul!`` / ``*`` / ``lu`` / ``ldiv!`` / ``\\``).

Julia is not available in this image, so the driver logic that sits above the native calls
is restated here in Python over the C ABI -- same names, argument meaning and error
behaviour as the reference, so that the parity tests read like the reference's own:

=====================================  ===========================================================
here                                   reference
=====================================  ===========================================================
``mul_(y, A, x, alpha, beta)``         ``mul!(y,A,x,α,β)`` -> ``_banded_muladd!`` src/generic/matmul.jl:41-64, 66-92
``mul_(C, A, B, alpha, beta)``         banded x banded ``gbmm!`` src/banded/gbmm.jl:207-293; banded x dense
                                       src/generic/matmul.jl:243-256; dense x banded :258-271
``matmul(A, B)``                       ``A*B`` (``similar(::MulAdd)`` src/generic/matmul.jl:1-6)
``lu(A)`` / ``lu_(A)``                 ``lu`` / ``lu!`` src/banded/BandedLU.jl:90-111
``ldiv_(F, B)`` / ``solve(A, b)``      ``ldiv!`` / ``\\`` src/banded/linalg.jl:5-9, 24-30, 41-47
=====================================  ===========================================================

Every arithmetic step -- including the band bookkeeping of the gbmm! driver (zero-band counts, block scaling, the transposed
copy) -- is a kernel of ``libbmb200.so``; nothing here computes on the CPU or through eager tensor ops.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .banded import (BandedMatrix, BandError, DimensionMismatch, LAPACKException, Transposed, bandwidths, colmajor)

vp = C.c_void_p


def _h(t: torch.Tensor) -> _lib.Handle:
    if not t.is_cuda:
        raise TypeError("operands must live on the GPU (no CPU fallback)")
    return _lib.handle(t.device.index if t.device.index is not None else torch.cuda.current_device())


def _inc(v: torch.Tensor) -> int:
    return int(v.stride(0)) if v.shape[0] > 1 else 1


def _ld(X: torch.Tensor) -> int:
    """Leading dimension of a column-major 2-D tensor (stride (1, ld))."""
    if X.shape[0] > 1 and X.stride(0) != 1:
        raise TypeError("dense matrices must be column-major (use colmajor()/to_colmajor())")
    return int(X.stride(1)) if X.shape[1] > 1 else max(1, int(X.shape[0]))


# ---------------------------------------------------------------------------------------------
# _fill_lmul! / _fill_rmul!  (src/generic/utils.jl:29-31): beta == 0 zero-fills (unless Czero)
# ---------------------------------------------------------------------------------------------
def _fill_vec(y: torch.Tensor, beta: float, yzero: bool = False) -> None:
    if y.numel() == 0 or (beta == 0 and yzero):
        return
    hd = _h(y)
    hd.check(hd.lib.bmb200_dfill_lmul(hd.h, float(beta), vp(y.data_ptr()), y.shape[0], 1, 0, _inc(y)), "dfill_lmul")


def _fill_cm(X: torch.Tensor, beta: float, zero: bool = False) -> None:
    """Column-major dense block (rows x cols, stride (1, ld))."""
    if X.numel() == 0 or (beta == 0 and zero):
        return
    hd = _h(X)
    hd.check(hd.lib.bmb200_dfill_lmul(hd.h, float(beta), vp(X.data_ptr()), X.shape[0], X.shape[1], _ld(X), 1),
             "dfill_lmul")


def _fill_banddata(data: torch.Tensor, beta: float, zero: bool = False) -> None:
    """Band-data block ``data[c0:c1, r0:r1]`` (tensor layout (n, rows), stride (lda, 1))."""
    if data.numel() == 0 or (beta == 0 and zero):
        return
    hd = _h(data)
    lda = int(data.stride(0)) if data.shape[0] > 1 else max(1, int(data.shape[1]))
    hd.check(hd.lib.bmb200_dfill_lmul(hd.h, float(beta), vp(data.data_ptr()), data.shape[1], data.shape[0], lda, 1),
             "dfill_lmul")


def _fill_banded_rows(Cm: BandedMatrix, r0: int, r1: int, c0: int, c1: int, beta: float, zero: bool) -> None:
    """lmul!(beta, view(C, r0+1:r1, c0+1:c1)) on a BandedMatrix: scale/zero the in-band entries of that block."""
    if r1 <= r0 or c1 <= c0 or Cm.data.numel() == 0 or (beta == 0 and zero):
        return
    hd = _h(Cm.data)
    hd.check(hd.lib.bmb200_dband_lmul_block(hd.h, Cm.m, Cm.n, Cm.l, Cm.u, vp(Cm.ptr), Cm.lda, r0, r1, c0, c1, float(beta)),
             "dband_lmul_block")


# ---------------------------------------------------------------------------------------------
# matrix * vector
# ---------------------------------------------------------------------------------------------
def _banded_gbmv(tA: str, alpha, A: BandedMatrix, x: torch.Tensor, beta, y: torch.Tensor, yzero=False):
    """_banded_gbmv! + banded_gbmv!  (src/generic/matmul.jl:21-39).  A has non-negative bandwidths here."""
    if y.shape[0] == 0:
        return y
    if x.shape[0] == 0:
        _fill_vec(y, beta, yzero)
        return y
    # Base.unalias(y, x): the reference copies x when it shares memory with y (matmul.jl:35)
    if x.untyped_storage().data_ptr() == y.untyped_storage().data_ptr():
        x = x.clone()
    hd = _h(y)
    hd.check(hd.lib.bmb200_dgbmv(hd.h, tA.encode(), A.m, A.n, A.l, A.u, float(alpha), vp(A.ptr), A.lda,
                                 vp(x.data_ptr()), _inc(x), float(beta), vp(y.data_ptr()), _inc(y)), "dgbmv")
    return y


def _banded_muladd_vec(alpha, A: BandedMatrix, x, beta, y, yzero=False):
    """_banded_muladd!(α,A,x,β,y)  (src/generic/matmul.jl:41-59)."""
    m, n = A.shape
    l, u = A.l, A.u
    if -l > u:
        _fill_vec(y, beta, yzero)
    elif l < 0:
        _banded_gbmv("N", alpha, A.view_cols(-l), x[-l:], beta, y, yzero)
    elif u < 0:
        _fill_vec(y[:-u], beta, yzero)
        _banded_gbmv("N", alpha, A.view_rows(-u), x, beta, y[-u:], yzero)
    else:
        _banded_gbmv("N", alpha, A, x, beta, y, yzero)
    return y


def _banded_muladd_row(alpha, At: BandedMatrix, x, beta, y, yzero=False):
    """_banded_muladd_row!('T', α, transpose(A), x, β, y)  (src/generic/matmul.jl:66-86).
    ``At`` is the PARENT (column-major) matrix of the transposed operand: y <- α At' x + β y."""
    n, m = At.shape
    u, l = At.l, At.u  # reference: u, l = bandwidths(At)
    if -l > u:
        _fill_vec(y, beta, yzero)
    elif l < 0:
        _banded_gbmv("T", alpha, At.view_rows(-l), x[-l:], beta, y, yzero)
    elif u < 0:
        _fill_vec(y[:-u], beta, yzero)
        _banded_gbmv("T", alpha, At.view_cols(-u), x, beta, y[-u:], yzero)
    else:
        _banded_gbmv("T", alpha, At, x, beta, y, yzero)
    return y


# ---------------------------------------------------------------------------------------------
# banded * banded : gbmm!  (src/banded/gbmm.jl:207-293)
# ---------------------------------------------------------------------------------------------
def _nonzero_band_rows(A: BandedMatrix) -> list:
    rows = A.l + A.u + 1
    if rows <= 0:
        return []
    hd = _h(A.data)
    flags = (C.c_int * rows)()
    hd.check(hd.lib.bmb200_dband_nonzero_rows(hd.h, A.m, A.n, A.l, A.u, vp(A.ptr), A.lda, flags), "dband_nonzero_rows")
    return list(flags)


def _num_zeroband_u(A: BandedMatrix) -> int:
    """gbmm.jl:191-197: number of leading all-zero upper bands (data rows from the top)."""
    f = _nonzero_band_rows(A)
    for b, nz in enumerate(f):
        if nz:
            return b
    return len(f)


def _num_zeroband_l(A: BandedMatrix) -> int:
    """gbmm.jl:199-205: number of trailing all-zero lower bands (data rows from the bottom)."""
    f = _nonzero_band_rows(A)
    for b, nz in enumerate(reversed(f)):
        if nz:
            return b
    return len(f)


def gbmm_(alpha, A: BandedMatrix, B: BandedMatrix, beta, Cm: BandedMatrix, Czero: bool = False):
    """gbmm!('N','N',α,A,B,β,C)  -- src/banded/gbmm.jl:207-293, branch for branch."""
    n, nu = A.shape
    m = B.n
    Am, An = A.shape
    Bm, Bn = B.shape
    assert n == Cm.m and nu == B.m and m == Cm.n
    if n == 0 or m == 0:
        return Cm
    Al, Au = A.l, A.u
    Bl, Bu = B.l, B.u
    Ctl, Ctu = Cm.l, Cm.u
    Cl, Cu = min(n - 1, Al + Bl), min(m - 1, Au + Bu)

    if (-Al > Au) or (-Bl > Bu):  # :231-233  A or B has no bands
        if not Czero:
            Cm.data.zero_()
        return Cm
    if Al < 0:  # :234-237
        _fill_banded_rows(Cm, max(1, Bn + Al - 1) - 1, Am, 0, m, beta, Czero)
        return gbmm_(alpha, A.view_cols(-Al), B.view_rows(-Al), beta, Cm, Czero)
    if Au < 0:  # :238-241
        _fill_banded_rows(Cm, 0, -Au, 0, m, beta, Czero)
        return_c = gbmm_(alpha, A.view_rows(-Au), B, beta, Cm.view_rows(-Au), Czero)
        del return_c
        return Cm
    if Bl < 0:  # :242-245
        _fill_banded_rows(Cm, 0, n, 0, -Bl, beta, Czero)
        gbmm_(alpha, A, B.view_cols(-Bl), beta, Cm.view_cols(-Bl), Czero)
        return Cm
    if Bu < 0:  # :246-249
        _fill_banded_rows(Cm, 0, n, max(1, Am + Bu - 1) - 1, Bn, beta, Czero)
        return gbmm_(alpha, A.view_cols(-Bu), B.view_rows(-Bu), beta, Cm, Czero)
    if Ctu < Cu:  # :250-264  C has too few upper bands
        Au_r, Bu_r = _num_zeroband_u(A), _num_zeroband_u(B)
        if not Ctu >= Cu - Au_r - Bu_r:
            raise BandError(Cm, Cu - Au_r - Bu_r)
        if Au - Au_r < -Al or Bu - Bu_r < -Bl:
            _fill_banddata(Cm.data, beta, Czero)
            return Cm
        At = BandedMatrix(A.data[:, Au_r:], n, Al, Au - Au_r)
        Bt = BandedMatrix(B.data[:, Bu_r:], nu, Bl, Bu - Bu_r)
        return gbmm_(alpha, At, Bt, beta, Cm, Czero)
    if Ctl < Cl:  # :265-280  too few lower bands
        Al_r, Bl_r = _num_zeroband_l(A), _num_zeroband_l(B)
        if not Ctl >= Cl - Al_r - Bl_r:
            raise BandError(Cm, Cl - Al_r - Bl_r)
        if Al - Al_r < -Au or Bl - Bl_r < -Bu:
            _fill_banddata(Cm.data, beta, Czero)
            return Cm
        At = BandedMatrix(A.data[:, : A.data.shape[1] - Al_r], n, Al - Al_r, Au)
        Bt = BandedMatrix(B.data[:, : B.data.shape[1] - Bl_r], nu, Bl - Bl_r, Bu)
        return gbmm_(alpha, At, Bt, beta, Cm, Czero)

    # :282-290  scale the extra bands of C, then hand the written band window to _gbmm!
    rowsC = Cm.data.shape[1]
    _fill_banddata(Cm.data[:, : min(Ctu - Cu, rowsC)], beta, Czero)
    _fill_banddata(Cm.data[:, Ctu + Cl + 1 :], beta, Czero)
    C_data = Cm.data[:, Ctu - Cu : Ctu + Cl + 1]
    _gbmm(alpha, A, B, beta, C_data, (n, nu, m), (Al, Au), (Bl, Bu), (Cl, Cu), Cm.lda)
    return Cm


def _gbmm(alpha, A, B, beta, C_data, sizes, Ab, Bb, Cb, ldc):
    """_gbmm!  (src/banded/gbmm.jl:296-340): one bmb200_dgbmm_bb launch instead of m dgbmv_ calls."""
    n, nu, m = sizes
    hd = _h(C_data)
    hd.check(hd.lib.bmb200_dgbmm_bb(hd.h, n, nu, m, Ab[0], Ab[1], Bb[0], Bb[1], Cb[0], Cb[1], float(alpha),
                                    vp(A.ptr), A.lda, vp(B.ptr), B.lda, float(beta), vp(C_data.data_ptr()), ldc),
             "dgbmm_bb")


# ---------------------------------------------------------------------------------------------
# banded * dense, dense * banded  (src/generic/matmul.jl:243-271)
# ---------------------------------------------------------------------------------------------
def _gbmm_bd(trans, alpha, A: BandedMatrix, B, beta, Cd):
    hd = _h(Cd)
    hd.check(hd.lib.bmb200_dgbmm_bd(hd.h, trans.encode(), A.m, A.n, A.l, A.u, Cd.shape[1], float(alpha), vp(A.ptr),
                                    A.lda, vp(B.data_ptr()), _ld(B), float(beta), vp(Cd.data_ptr()), _ld(Cd)),
             "dgbmm_bd")


def _banded_times_dense(alpha, A, B, beta, Cd):
    """materialize!(MatMulMatAdd{BandedColumns,Strided,Strided}) matmul.jl:243-256: the reference loops
    mul!(colC, A, colB, α, β) over columns; all columns go through one multi-RHS launch here, with the
    same negative-bandwidth re-viewing as _banded_muladd! (matmul.jl:41-59)."""
    if alpha == 0:
        _fill_cm(Cd, beta)
        return Cd
    tr = isinstance(A, Transposed)
    P = A.parent if tr else A
    if Cd.shape[0] == 0 or Cd.shape[1] == 0:
        return Cd
    if B.shape[0] == 0:
        _fill_cm(Cd, beta)
        return Cd
    if not tr:
        l, u = P.l, P.u
        if -l > u:
            _fill_cm(Cd, beta)
        elif l < 0:
            _gbmm_bd("N", alpha, P.view_cols(-l), B[-l:], beta, Cd)
        elif u < 0:
            _fill_cm(Cd[:-u], beta)
            _gbmm_bd("N", alpha, P.view_rows(-u), B, beta, Cd[-u:])
        else:
            _gbmm_bd("N", alpha, P, B, beta, Cd)
    else:
        u, l = P.l, P.u
        if -l > u:
            _fill_cm(Cd, beta)
        elif l < 0:
            _gbmm_bd("T", alpha, P.view_rows(-l), B[-l:], beta, Cd)
        elif u < 0:
            _fill_cm(Cd[:-u], beta)
            _gbmm_bd("T", alpha, P.view_cols(-u), B, beta, Cd[-u:])
        else:
            _gbmm_bd("T", alpha, P, B, beta, Cd)
    return Cd


def _dense_times_banded(alpha, Ad, B, beta, Cd):
    """matmul.jl:258-271: for each row, mul!(rowC, transpose(B), rowA, α, β) (strided x and y).  One bmb200_dgbmm_db launch
    covers all rows when the band widths are non-negative; the negative-bandwidth re-viewing of _banded_muladd!
    (matmul.jl:41-59, 66-86) keeps the reference's per-row form."""
    if alpha == 0:
        _fill_cm(Cd, beta)
        return Cd
    tr = isinstance(B, Transposed)
    P = B.parent if tr else B
    if Cd.shape[0] == 0 or Cd.shape[1] == 0:
        return Cd
    if P.l >= 0 and P.u >= 0 and Ad.shape[1] > 0:
        hd = _h(Cd)
        hd.check(hd.lib.bmb200_dgbmm_db(hd.h, (b"T" if tr else b"N"), Cd.shape[0], Ad.shape[1], Cd.shape[1], P.l, P.u,
                                        float(alpha), vp(Ad.data_ptr()), _ld(Ad), vp(P.ptr), P.lda, float(beta),
                                        vp(Cd.data_ptr()), _ld(Cd)), "dgbmm_db")
        return Cd
    for i in range(Cd.shape[0]):
        if tr:
            _banded_muladd_vec(alpha, P, Ad[i], beta, Cd[i])
        else:
            _banded_muladd_row(alpha, P, Ad[i], beta, Cd[i])
    return Cd


# ---------------------------------------------------------------------------------------------
# public: mul!, *
# ---------------------------------------------------------------------------------------------
def _shape(X):
    return tuple(X.shape)


def mul_(Cout, A, B, alpha=1.0, beta=0.0):
    """``mul!(C, A, B, α, β)``: C <- α A B + β C, dispatching like the reference's MulAdd layouts."""
    sa, sb, sc = _shape(A), _shape(B), _shape(Cout)
    if len(sb) == 1:  # matrix * vector (checkdimensions: DimensionMismatch)
        if sa[1] != sb[0] or sa[0] != sc[0] or len(sc) != 1:
            raise DimensionMismatch(f"A has dimensions {sa} but B has dimensions {sb} and C {sc}")
        if isinstance(A, Transposed):
            return _banded_muladd_row(alpha, A.parent, B, beta, Cout)
        return _banded_muladd_vec(alpha, A, B, beta, Cout)
    if sa[1] != sb[0] or sc != (sa[0], sb[1]):
        raise DimensionMismatch(f"A has dimensions {sa} but B has dimensions {sb} and C {sc}")
    a_b = isinstance(A, (BandedMatrix, Transposed))
    b_b = isinstance(B, (BandedMatrix, Transposed))
    if a_b and b_b:
        if not isinstance(Cout, BandedMatrix):
            raise TypeError("banded*banded needs a BandedMatrix destination")
        # matmul.jl:182-184: non column-major operands are first converted to plain BandedMatrix
        A2 = materialize_transpose(A) if isinstance(A, Transposed) else A
        B2 = materialize_transpose(B) if isinstance(B, Transposed) else B
        return gbmm_(alpha, A2, B2, beta, Cout)
    if a_b:
        return _banded_times_dense(alpha, A, B, beta, Cout)
    if b_b:
        return _dense_times_banded(alpha, A, B, beta, Cout)
    raise TypeError("at least one operand must be banded")


def materialize_transpose(At: Transposed) -> BandedMatrix:
    """convert(DefaultBandedMatrix, A') (matmul.jl:182-184): band row r of A' is band row (l+u-r) of A, shifted."""
    P = At.parent
    out = BandedMatrix.undef((P.n, P.m), (P.u, P.l), device=P.data.device)
    if P.l + P.u + 1 > 0 and P.m > 0 and P.n > 0:
        hd = _h(P.data)
        hd.check(hd.lib.bmb200_dband_transpose(hd.h, P.m, P.n, P.l, P.u, vp(P.ptr), P.lda, vp(out.ptr), out.lda), "dband_transpose")
    return out


def matmul(A, B):
    """``A*B``: allocates the destination like similar(::MulAdd) (src/generic/matmul.jl:1-6)."""
    sa, sb = _shape(A), _shape(B)
    a_b = isinstance(A, (BandedMatrix, Transposed))
    b_b = isinstance(B, (BandedMatrix, Transposed))
    dev = (A.parent if isinstance(A, Transposed) else A).data.device if a_b else (
        B.parent if isinstance(B, Transposed) else B).data.device
    if len(sb) == 1:
        if sa[1] != sb[0]:
            raise DimensionMismatch(f"second dimension of A, {sa[1]}, does not match length of x, {sb[0]}")
        y = torch.empty(sa[0], dtype=torch.float64, device=dev)
        return mul_(y, A, B, 1.0, 0.0)
    if sa[1] != sb[0]:
        raise DimensionMismatch(f"A has dimensions {sa} but B has dimensions {sb}")
    if a_b and b_b:
        (Al, Au), (Bl, Bu) = bandwidths(A), bandwidths(B)
        bw = (min(sa[0] - 1, Al + Bl), min(sb[1] - 1, Au + Bu))  # bandwidths(M) = min.(_bnds(M), prodbandwidths)
        Cm = BandedMatrix.undef((sa[0], sb[1]), bw, device=dev)
        return mul_(Cm, A, B, 1.0, 0.0)
    Cd = colmajor(sa[0], sb[1], device=dev)
    return mul_(Cd, A, B, 1.0, 0.0)


# ---------------------------------------------------------------------------------------------
# lu / ldiv! / \
# ---------------------------------------------------------------------------------------------
class BandedLU:
    """``BandedLU{T,S}`` (src/banded/BandedLU.jl:10-19): factors (bandwidths (l, l+u)), HOST ipiv, info."""

    def __init__(self, factors: BandedMatrix, ipiv: np.ndarray, info: int, d_ipiv: torch.Tensor | None = None):
        self.factors, self.ipiv, self.info = factors, ipiv, int(info)
        self._d_ipiv = d_ipiv

    @property
    def shape(self):
        return self.factors.shape

    @property
    def T(self):
        return TransposeFact(self)

    def issuccess(self) -> bool:
        return self.info == 0

    @property
    def p(self) -> np.ndarray:
        """ipiv2perm (BandedLU.jl:125), 1-based like the reference."""
        m = self.factors.m
        p = np.arange(1, m + 1)
        for i, piv in enumerate(self.ipiv):
            p[i], p[piv - 1] = p[piv - 1], p[i]
        return p

    def d_ipiv(self) -> torch.Tensor:
        if self._d_ipiv is None:
            self._d_ipiv = torch.as_tensor(self.ipiv).to(self.factors.data.device)
        return self._d_ipiv


class TransposeFact:
    def __init__(self, parent: BandedLU):
        self.parent = parent


def lu_(A: BandedMatrix, check: bool = True) -> BandedLU:
    """``lu!(A)`` (BandedLU.jl:90-103): A has bandwidths (l, l+u_orig) with zeros in the extra l bands."""
    m = A.m
    l, u = A.l, A.u
    if m == 0:
        return BandedLU(A, np.zeros(0, dtype=np.int64), 0)
    hd = _h(A.data)
    mn = min(m, A.n)
    d_ipiv = torch.empty(mn, dtype=torch.int64, device=A.data.device)
    info = C.c_int(0)
    rc = hd.lib.bmb200_dgbtrf(hd.h, m, A.n, l, u - l, vp(A.ptr), A.lda, vp(d_ipiv.data_ptr()), C.byref(info))
    if rc < 0 and rc > -100:
        raise ValueError(f"invalid argument #{-rc} to LAPACK call")  # chklapackerror(info<0) -> ArgumentError
    hd.check(rc, "dgbtrf")
    if info.value > 0:
        raise LAPACKException(info.value)  # gbtrf! -> chklapackerror, BandedLU.jl:98
    return BandedLU(A, d_ipiv.cpu().numpy(), 0, d_ipiv)


def lu(A: BandedMatrix, check: bool = True) -> BandedLU:
    """``lu(A)`` -> ``_lu`` (BandedLU.jl:108-111): widening copy into (l, l+u) storage, then lu!."""
    l, u = A.l, A.u
    if A.m == 0 or A.n == 0:  # zero-size (test_bandedlu.jl:157-164): factors are zeros, no LAPACK call for m == 0
        W = BandedMatrix.zeros(A.shape, (l, l + u), device=A.data.device)
        return lu_(W, check)
    if l < 0 or l + u + 1 <= 0:
        raise ValueError("invalid argument #1 to LAPACK call")  # gbtrf!(kl<0, ...) -> ArgumentError in the reference
    W = BandedMatrix.undef(A.shape, (l, l + u), device=A.data.device)
    hd = _h(A.data)
    mn = min(A.m, A.n)
    d_ipiv = torch.empty(mn, dtype=torch.int64, device=A.data.device)
    info = C.c_int(0)
    rc = hd.lib.bmb200_dgbtrf_from(hd.h, A.m, A.n, l, u, vp(A.ptr), A.lda, vp(W.ptr), W.lda, vp(d_ipiv.data_ptr()), C.byref(info))
    if rc < 0 and rc > -100:
        raise ValueError(f"invalid argument #{-rc} to LAPACK call")
    hd.check(rc, "dgbtrf_from")
    if info.value > 0:
        raise LAPACKException(info.value)  # gbtrf! -> chklapackerror, BandedLU.jl:98
    return BandedLU(W, d_ipiv.cpu().numpy(), 0, d_ipiv)


def ldiv_(F, B: torch.Tensor) -> torch.Tensor:
    """``ldiv!(F, B)`` (linalg.jl:24-30) and ``ldiv!(transpose(F), B)`` (:41-47); B is overwritten."""
    trans = "N"
    if isinstance(F, TransposeFact):
        F, trans = F.parent, "T"
    A = F.factors
    m = A.m
    if B.shape[0] != m:
        raise DimensionMismatch(f"B has first dimension {B.shape[0]} but needs {m}")
    if m == 0:
        return B
    l, u = A.l, A.u
    hd = _h(B)
    nrhs = 1 if B.dim() == 1 else B.shape[1]
    if B.dim() == 1 and _inc(B) != 1:
        raise TypeError("right-hand side vector must be contiguous")
    ldb = max(1, m) if B.dim() == 1 else _ld(B)
    rc = hd.lib.bmb200_dgbtrs(hd.h, trans.encode(), m, l, u - l, nrhs, vp(A.ptr), A.lda, vp(F.d_ipiv().data_ptr()),
                              vp(B.data_ptr()), ldb)
    hd.check(rc, "dgbtrs")
    return B


# ---------------------------------------------------------------------------------------------------
# Triangular band solve / multiply: tbsv! / tbmv! (src/blas.jl:71-141) and the UpperTriangular / LowerTriangular
# {<:BandedMatrix} ldiv! / lmul! that reach them (src/tribanded.jl:47-84)
# ---------------------------------------------------------------------------------------------------
def _tb(fn_name: str, uplo: str, trans: str, diag: str, m: int, k: int, Adata: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """``Adata`` is the band array as this package stores it: shape (n, rows >= k+1), one column of the band per row of the
    tensor, ``Adata.stride(0)`` = lda (so a row-range view of ``BandedMatrix.data`` passes straight through)."""
    n = Adata.shape[0]
    if Adata.shape[1] < k + 1:  # blas.jl:88,126
        raise ValueError("triangular banded data missing")
    if n != m:
        raise DimensionMismatch(f"matrix is not square: dimensions are {n}, {m}")
    if n != x.shape[0]:
        raise DimensionMismatch(f"size of A is {n} != length(x) = {x.shape[0]}")
    if x.dim() != 1 or _inc(x) != 1:
        raise TypeError("x must be a contiguous vector")
    if Adata.shape[1] > 1 and Adata.stride(1) != 1:  # chkstride1
        raise ValueError("band rows of one column must be contiguous")
    if n == 0:
        return x
    hd = _h(x)
    lda = max(1, Adata.stride(0)) if n > 1 else max(1, Adata.shape[1])
    p = _typed(Adata, x)  # Float64 keeps its tuned kernels; S / C / Z run the generic ones (same argument list)
    fn_name = fn_name.replace("bmb200_d", f"bmb200_{p}")
    rc = getattr(hd.lib, fn_name)(hd.h, uplo.encode(), trans.encode(), diag.encode(), n, k, vp(Adata.data_ptr()), lda,
                                  vp(x.data_ptr()), 1)
    hd.check(rc, fn_name[7:])
    return x


def tbsv_(uplo, trans, diag, m, k, Adata, x):
    """``tbsv!(uplo, trans, diag, m, k, A, x)`` (src/blas.jl:121-141): x <- inv(T) x in place."""
    return _tb("bmb200_dtbsv", uplo, trans, diag, m, k, Adata, x)


def tbmv_(uplo, trans, diag, m, k, Adata, x):
    """``tbmv!(uplo, trans, diag, m, k, A, x)`` (src/blas.jl:83-101): x <- T x in place."""
    return _tb("bmb200_dtbmv", uplo, trans, diag, m, k, Adata, x)


def _tri_data(uplo: str, A: BandedMatrix):
    """bandeddata of the triangular view: rows 1:u+1 of A.data for 'U', rows u+1:u+l+1 for 'L'; bandwidth(A, bwdim)."""
    if A.m != A.n:
        raise DimensionMismatch(f"matrix is not square: dimensions are {A.shape}")
    if uplo == "U":
        return A.data[:, : A.u + 1], A.u
    return A.data[:, A.u:], A.l


def ldiv_tri_(uplo: str, unit: bool, A: BandedMatrix, x: torch.Tensor) -> torch.Tensor:
    """``ldiv!(UpperTriangular(A), x)`` / LowerTriangular / Unit* (src/tribanded.jl:75-84)."""
    d, k = _tri_data(uplo, A)
    return tbsv_(uplo, "N", "U" if unit else "N", A.m, k, d, x)


def lmul_tri_(uplo: str, unit: bool, A: BandedMatrix, x: torch.Tensor) -> torch.Tensor:
    """``lmul!(UpperTriangular(A), x)`` / LowerTriangular / Unit* (src/tribanded.jl:47-55)."""
    d, k = _tri_data(uplo, A)
    return tbmv_(uplo, "N", "U" if unit else "N", A.m, k, d, x)


# ---------------------------------------------------------------------------------------------------
# Symmetric band matvec: sbmv! (src/blas.jl:36-66) and mul! of Symmetric{<:BandedMatrix} (src/symbanded/symbanded.jl:72-93)
# ---------------------------------------------------------------------------------------------------
def sbmv_(uplo: str, k: int, alpha, Adata: torch.Tensor, x: torch.Tensor, beta, y: torch.Tensor) -> torch.Tensor:
    """``sbmv!(uplo, k, alpha, A, x, beta, y)`` (src/blas.jl:64-66); ``Adata`` is the (n, rows >= k+1) band array of the stored
    triangle (a row-range view of ``BandedMatrix.data``)."""
    n = Adata.shape[0]
    if x.shape[0] != n or y.shape[0] != n:
        raise DimensionMismatch("*")
    if Adata.shape[1] < k + 1:
        raise ValueError("symmetric banded data missing")
    if x.dim() != 1 or y.dim() != 1 or _inc(x) != 1 or _inc(y) != 1:
        raise TypeError("x and y must be contiguous vectors")
    if n == 0:
        return y
    hd = _h(y)
    lda = max(1, Adata.stride(0)) if n > 1 else max(1, Adata.shape[1])
    rc = hd.lib.bmb200_dsbmv(hd.h, uplo.encode(), n, k, float(alpha), vp(Adata.data_ptr()), lda, vp(x.data_ptr()), 1, float(beta),
                             vp(y.data_ptr()), 1)
    hd.check(rc, "dsbmv")
    return y


def mul_sym_(y: torch.Tensor, uplo: str, A: BandedMatrix, x: torch.Tensor, alpha=1.0, beta=0.0) -> torch.Tensor:
    """``mul!(y, Symmetric(A, uplo), x, alpha, beta)`` (symbanded.jl:84-93): only A's `uplo` triangle is read."""
    if A.m != A.n:
        raise DimensionMismatch("matrix is not square")
    if y.shape[0] != A.m or x.shape[0] != A.n:
        raise DimensionMismatch("*")
    k = A.u if uplo == "U" else A.l  # bandwidth(Symmetric(A, uplo))
    if k < 0:
        _fill_vec(y, beta)
        return y
    if x.data_ptr() == y.data_ptr():  # _banded_sbmv!: x === y -> copy(x)
        x = x.clone()
    d, _ = _tri_data(uplo, A)
    return sbmv_(uplo, k, alpha, d, x, beta, y)


# ---------------------------------------------------------------------------------------------------
# Banded Cholesky: pbtrf! / pbtrs! (src/lapack.jl:268-332), cholesky / cholesky! of Symmetric{<:BandedMatrix}
# (banded_chol!, src/symbanded/BandedCholesky.jl:2-13) and ldiv! of the factorisation (:72-80)
# ---------------------------------------------------------------------------------------------------
class PosDefException(Exception):
    """LinearAlgebra.PosDefException(info): the leading minor of order ``info`` is not positive definite."""

    def __init__(self, info):
        super().__init__(f"matrix is not positive definite; Cholesky factorization failed (leading minor of order {info}).")
        self.info = int(info)


def pbtrf_(uplo: str, m: int, kd: int, Adata: torch.Tensor):
    """``pbtrf!(uplo, m, kd, A)`` (src/lapack.jl:275-291): A is the (n, rows >= kd+1) band array of the stored triangle
    (a row-range view of ``BandedMatrix.data``), factored in place.  Returns (Adata, info)."""
    if uplo not in ("U", "L"):  # chkuplo
        raise ValueError(f"uplo argument must be 'U' (upper) or 'L' (lower), got {uplo}")
    n = Adata.shape[0]
    if n != m:
        raise ValueError("Matrix must be square")  # lapack.jl:281
    if Adata.shape[1] < kd + 1:
        raise ValueError("Not enough bands")  # lapack.jl:282
    if Adata.shape[1] > 1 and Adata.stride(1) != 1:  # chkstride1
        raise ValueError("band rows of one column must be contiguous")
    if n == 0:
        return Adata, 0
    hd = _h(Adata)
    lda = max(1, Adata.stride(0)) if n > 1 else max(1, Adata.shape[1])
    info = C.c_int(0)
    rc = getattr(hd.lib, f"bmb200_{_typed(Adata)}pbtrf")(hd.h, uplo.encode(), n, kd, vp(Adata.data_ptr()), lda, C.byref(info))
    if rc < 0 and rc > -100:
        raise ValueError(f"invalid argument #{-rc} to LAPACK call")  # chkargsok
    hd.check(rc, "dpbtrf")
    return Adata, info.value


def pbtrs_(uplo: str, m: int, kd: int, Adata: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
    """``pbtrs!(uplo, m, kd, A, B)`` (src/lapack.jl:305-329): B <- inv(A) B from the factor held in ``Adata``."""
    if uplo not in ("U", "L"):
        raise ValueError(f"uplo argument must be 'U' (upper) or 'L' (lower), got {uplo}")
    n = Adata.shape[0]
    if m != n or m != B.shape[0]:  # lapack.jl:311-313
        raise DimensionMismatch(f"matrix A has dimensions ({n}, {n}), but right hand side matrix B has dimensions {tuple(B.shape)}")
    if Adata.shape[1] < kd + 1:
        raise ValueError("Not enough bands")
    if n == 0:
        return B
    if B.dim() == 1 and _inc(B) != 1:
        raise TypeError("right-hand side vector must be contiguous")
    nrhs = 1 if B.dim() == 1 else B.shape[1]
    ldb = max(1, n) if B.dim() == 1 else _ld(B)
    hd = _h(B)
    lda = max(1, Adata.stride(0)) if n > 1 else max(1, Adata.shape[1])
    rc = getattr(hd.lib, f"bmb200_{_typed(Adata, B)}pbtrs")(hd.h, uplo.encode(), n, kd, nrhs, vp(Adata.data_ptr()), lda, vp(B.data_ptr()), ldb)
    if rc < 0 and rc > -100:
        raise ValueError(f"invalid argument #{-rc} to LAPACK call")
    hd.check(rc, "dpbtrs")
    return B


class BandedCholesky:
    """``Cholesky{T,<:BandedMatrix}``: ``factors`` holds U (uplo 'U': A = U'U, bandwidths (., u)) or L ('L') in the stored
    triangle of a BandedMatrix; the other triangle is not referenced."""

    def __init__(self, factors: BandedMatrix, uplo: str, info: int):
        self.factors, self.uplo, self.info = factors, uplo, int(info)

    @property
    def shape(self):
        return self.factors.shape

    def issuccess(self) -> bool:
        return self.info == 0

    def _tri(self):
        return _tri_data(self.uplo, self.factors)


def cholesky_(A: BandedMatrix, uplo: str = "U", check: bool = True) -> BandedCholesky:
    """``cholesky!(Symmetric(A, uplo))`` -> banded_chol! (BandedCholesky.jl:2-13): the `uplo` triangle of A is overwritten."""
    d, k = _tri_data(uplo, A)
    _, info = pbtrf_(uplo, A.m, max(k, 0), d)
    if check and info != 0:
        raise PosDefException(info)  # checkpositivedefinite
    return BandedCholesky(A, uplo, info)


def cholesky(A: BandedMatrix, uplo: str = "U", check: bool = True) -> BandedCholesky:
    """``cholesky(Symmetric(A, uplo))`` (BandedCholesky.jl:83-89: cholcopy, then cholesky!)."""
    return cholesky_(A.copy(), uplo, check)


def ldiv_chol_(F: BandedCholesky, B: torch.Tensor) -> torch.Tensor:
    """``ldiv!(F::Cholesky{T,<:BandedMatrix}, B)`` (BandedCholesky.jl:62-80) -> pbtrs!; B is overwritten."""
    d, k = F._tri()
    return pbtrs_(F.uplo, F.factors.m, max(k, 0), d, B)


# ---------------------------------------------------------------------------------------------------
# The four BLAS element types: gbmv! / sbmv! / hbmv! (src/blas.jl:4-66) and LAPACK.gbtrf! / gbtrs! as the reference's lu /
# ldiv! call them (BandedLU.jl:90-103, linalg.jl:24-63) for Float32, Float64, ComplexF32, ComplexF64.  Band arrays are the
# package's (n, rows) tensors (one band column per tensor row), dtype float32 / float64 / complex64 / complex128.
# ---------------------------------------------------------------------------------------------------
_PREFIX = {torch.float32: "s", torch.float64: "d", torch.complex64: "c", torch.complex128: "z"}
_NPTYPE = {torch.float32: np.float32, torch.float64: np.float64, torch.complex64: np.complex64, torch.complex128: np.complex128}


def _typed(t: torch.Tensor, *others: torch.Tensor) -> str:
    if t.dtype not in _PREFIX:
        raise TypeError(f"element type {t.dtype} is not a BLAS float (Float32, Float64, ComplexF32, ComplexF64)")
    for o in others:
        if o.dtype != t.dtype:
            raise TypeError(f"element types differ: {t.dtype} and {o.dtype}")
    return _PREFIX[t.dtype]


def _host_scalar(dtype: torch.dtype, v) -> np.ndarray:
    return np.array([v], dtype=_NPTYPE[dtype])


def _lda_of(Adata: torch.Tensor) -> int:
    return max(1, Adata.stride(0)) if Adata.shape[0] > 1 else max(1, Adata.shape[1])


def gbmv_(trans: str, m: int, kl: int, ku: int, alpha, Adata: torch.Tensor, x: torch.Tensor, beta, y: torch.Tensor) -> torch.Tensor:
    """``gbmv!(trans, m, kl, ku, alpha, A, x, beta, y)`` (src/blas.jl:30-32) for the four element types; 'C' conjugates."""
    p = _typed(Adata, x, y)
    n = Adata.shape[0]
    if Adata.shape[1] < kl + ku + 1:
        raise ValueError("band data has fewer than kl+ku+1 rows")
    lenx, leny = (n, m) if trans == "N" else (m, n)
    if x.shape[0] != lenx or y.shape[0] != leny:
        raise DimensionMismatch("*")
    if m == 0 or n == 0:
        return y
    hd = _h(y)
    lda = _lda_of(Adata)
    if p == "d":
        rc = hd.lib.bmb200_dgbmv(hd.h, trans.encode(), m, n, kl, ku, float(alpha), vp(Adata.data_ptr()), lda, vp(x.data_ptr()), _inc(x),
                                 float(beta), vp(y.data_ptr()), _inc(y))
    else:
        al, be = _host_scalar(Adata.dtype, alpha), _host_scalar(Adata.dtype, beta)
        rc = getattr(hd.lib, f"bmb200_{p}gbmv")(hd.h, trans.encode(), m, n, kl, ku, vp(al.ctypes.data), vp(Adata.data_ptr()), lda,
                                                vp(x.data_ptr()), _inc(x), vp(be.ctypes.data), vp(y.data_ptr()), _inc(y))
    hd.check(rc, f"{p}gbmv")
    return y


def hbmv_(uplo: str, k: int, alpha, Adata: torch.Tensor, x: torch.Tensor, beta, y: torch.Tensor) -> torch.Tensor:
    """``sbmv!`` / ``hbmv!(uplo, k, alpha, A, x, beta, y)`` (src/blas.jl:36-66): the Hermitian (real types: symmetric) band
    matvec from the stored ``uplo`` triangle; mul! of Symmetric / Hermitian{<:BandedMatrix} (symbanded.jl:72-96)."""
    p = _typed(Adata, x, y)
    if p == "d":
        return sbmv_(uplo, k, alpha, Adata, x, beta, y)
    n = Adata.shape[0]
    if x.shape[0] != n or y.shape[0] != n:
        raise DimensionMismatch("*")
    if Adata.shape[1] < k + 1:
        raise ValueError("symmetric banded data missing")
    if n == 0:
        return y
    if x.data_ptr() == y.data_ptr():
        x = x.clone()
    hd = _h(y)
    al, be = _host_scalar(Adata.dtype, alpha), _host_scalar(Adata.dtype, beta)
    name = "bmb200_ssbmv" if p == "s" else f"bmb200_{p}hbmv"
    rc = getattr(hd.lib, name)(hd.h, uplo.encode(), n, k, vp(al.ctypes.data), vp(Adata.data_ptr()), _lda_of(Adata), vp(x.data_ptr()), 1,
                               vp(be.ctypes.data), vp(y.data_ptr()), 1)
    hd.check(rc, name[7:])
    return y


def gbmm_typed_(alpha, Adata: torch.Tensor, Bdata: torch.Tensor, beta, Cdata: torch.Tensor, sizes, Ab, Bb, Cb) -> torch.Tensor:
    """``_gbmm!(alpha, A_data, B_data, beta, C_data, (n, nu, m), (Al, Au), (Bl, Bu), (Cl, Cu))`` (src/banded/gbmm.jl:296-340) on
    band arrays of any of the four element types ((ncols, rows) tensors)."""
    p = _typed(Adata, Bdata, Cdata)
    n, nu, m = sizes
    hd = _h(Cdata)
    if p == "d":
        rc = hd.lib.bmb200_dgbmm_bb(hd.h, n, nu, m, Ab[0], Ab[1], Bb[0], Bb[1], Cb[0], Cb[1], float(alpha), vp(Adata.data_ptr()), _lda_of(Adata),
                                    vp(Bdata.data_ptr()), _lda_of(Bdata), float(beta), vp(Cdata.data_ptr()), _lda_of(Cdata))
    else:
        al, be = _host_scalar(Adata.dtype, alpha), _host_scalar(Adata.dtype, beta)
        rc = getattr(hd.lib, f"bmb200_{p}gbmm_bb")(hd.h, n, nu, m, Ab[0], Ab[1], Bb[0], Bb[1], Cb[0], Cb[1], vp(al.ctypes.data), vp(Adata.data_ptr()),
                                                   _lda_of(Adata), vp(Bdata.data_ptr()), _lda_of(Bdata), vp(be.ctypes.data), vp(Cdata.data_ptr()),
                                                   _lda_of(Cdata))
    hd.check(rc, f"{p}gbmm_bb")
    return Cdata


def gbmm_bd_typed_(trans: str, m: int, kl: int, ku: int, alpha, Adata: torch.Tensor, B: torch.Tensor, beta, Cd: torch.Tensor) -> torch.Tensor:
    """C <- alpha*op(A)*B + beta*C, A banded m x n, B / C dense column-major (the per-column mul! loop of
    src/generic/matmul.jl:243-256) for any of the four element types."""
    p = _typed(Adata, B, Cd)
    n = Adata.shape[0]
    hd = _h(Cd)
    nrhs = B.shape[1]
    if p == "d":
        rc = hd.lib.bmb200_dgbmm_bd(hd.h, trans.encode(), m, n, kl, ku, nrhs, float(alpha), vp(Adata.data_ptr()), _lda_of(Adata), vp(B.data_ptr()),
                                    _ld(B), float(beta), vp(Cd.data_ptr()), _ld(Cd))
    else:
        al, be = _host_scalar(Adata.dtype, alpha), _host_scalar(Adata.dtype, beta)
        rc = getattr(hd.lib, f"bmb200_{p}gbmm_bd")(hd.h, trans.encode(), m, n, kl, ku, nrhs, vp(al.ctypes.data), vp(Adata.data_ptr()), _lda_of(Adata),
                                                   vp(B.data_ptr()), _ld(B), vp(be.ctypes.data), vp(Cd.data_ptr()), _ld(Cd))
    hd.check(rc, f"{p}gbmm_bd")
    return Cd


def gbtrf_(m: int, kl: int, ku: int, AB: torch.Tensor):
    """``LAPACK.gbtrf!(kl, ku, m, AB)`` (BandedLU.jl:98) for the four element types: AB is the (n, rows >= 2kl+ku+1) LU-storage
    tensor, factored in place.  Returns (AB, d_ipiv [device int64, 1-based], info)."""
    p = _typed(AB)
    n = AB.shape[0]
    mn = min(m, n)
    d_ipiv = torch.empty(mn, dtype=torch.int64, device=AB.device)
    if mn == 0:
        return AB, d_ipiv, 0
    hd = _h(AB)
    info = C.c_int(0)
    rc = getattr(hd.lib, f"bmb200_{p}gbtrf")(hd.h, m, n, kl, ku, vp(AB.data_ptr()), _lda_of(AB), vp(d_ipiv.data_ptr()), C.byref(info))
    if rc < 0 and rc > -100:
        raise ValueError(f"invalid argument #{-rc} to LAPACK call")
    hd.check(rc, f"{p}gbtrf")
    return AB, d_ipiv, info.value


def gbtrs_(trans: str, kl: int, ku: int, m: int, AB: torch.Tensor, d_ipiv: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
    """``LAPACK.gbtrs!(trans, kl, ku, m, AB, ipiv, B)`` (linalg.jl:28,46,62) for the four element types; 'C' is the
    conjugate-transpose solve.  B (vector or column-major matrix) is overwritten."""
    p = _typed(AB, B)
    n = AB.shape[0]
    if B.shape[0] != n:
        raise DimensionMismatch(f"B has first dimension {B.shape[0]} but needs {n}")
    if n == 0:
        return B
    nrhs = 1 if B.dim() == 1 else B.shape[1]
    ldb = max(1, n) if B.dim() == 1 else _ld(B)
    hd = _h(B)
    rc = getattr(hd.lib, f"bmb200_{p}gbtrs")(hd.h, trans.encode(), n, kl, ku, nrhs, vp(AB.data_ptr()), _lda_of(AB), vp(d_ipiv.data_ptr()),
                                             vp(B.data_ptr()), ldb)
    if rc < 0 and rc > -100:
        raise ValueError(f"invalid argument #{-rc} to LAPACK call")
    hd.check(rc, f"{p}gbtrs")
    return B


# ---------------------------------------------------------------------------------------------------
# Band-aligned elementwise operations between different bandwidths (the steps either side of the hot path):
# banded_axpy! (src/banded/BandedMatrix.jl:1006-1015, src/generic/broadcast.jl:978-1020) and copyto! (broadcast.jl:175-230)
# ---------------------------------------------------------------------------------------------------
def _ld_band(A: BandedMatrix) -> int:
    return max(1, A.data.stride(0)) if A.n > 1 else max(1, A.data.shape[1])


def axpy_(a, X: BandedMatrix, Y: BandedMatrix) -> BandedMatrix:
    """``axpy!(a, X, Y)``: Y += a*X.  Equal bandwidths: one FMA per slot of the data arrays (``axpy!(a, X.data, Y.data)``);
    otherwise ``a*X[k,j] + Y[k,j]`` on the overlapping bands, ``BandError`` if X has a non-zero entry outside Y's bands."""
    if X.shape != Y.shape:
        raise DimensionMismatch(f"X has size {X.shape} but Y has size {Y.shape}")
    if X.m == 0 or X.n == 0:
        return Y
    hd = _h(Y.data)
    out = C.c_int64(0)
    rc = hd.lib.bmb200_dband_axpy(hd.h, X.m, X.n, float(a), X.l, X.u, vp(X.data.data_ptr()), _ld_band(X), Y.l, Y.u,
                                  vp(Y.data.data_ptr()), _ld_band(Y), C.byref(out))
    hd.check(rc, "dband_axpy")
    if out.value:
        raise BandError(Y, (X.l if X.l > Y.l else -X.u))
    return Y


def copyto_(dest: BandedMatrix, src: BandedMatrix) -> BandedMatrix:
    """``copyto!(dest, src)`` between bandwidths: overlapping bands copied, dest's other bands zeroed, ``BandError`` if src has
    a non-zero entry outside dest's bands."""
    if dest.shape != src.shape:
        raise DimensionMismatch(f"dest has size {dest.shape} but src has size {src.shape}")
    if dest.m == 0 or dest.n == 0:
        return dest
    hd = _h(dest.data)
    out = C.c_int64(0)
    rc = hd.lib.bmb200_dband_copy(hd.h, src.m, src.n, src.l, src.u, vp(src.data.data_ptr()), _ld_band(src), dest.l, dest.u,
                                  vp(dest.data.data_ptr()), _ld_band(dest), C.byref(out))
    hd.check(rc, "dband_copy")
    if out.value:
        raise BandError(dest, (src.l if src.l > dest.l else -src.u))
    return dest


def similar(A: BandedMatrix, bandwidths=None) -> BandedMatrix:
    """``similar(A)`` / ``similar(A, T, m, n, l, u)``: an uninitialised device BandedMatrix of the same size."""
    bw = (A.l, A.u) if bandwidths is None else tuple(bandwidths)
    return BandedMatrix.undef(A.shape, bw, device=A.data.device)


def axpby_(alpha, X: BandedMatrix, beta, Y: BandedMatrix, Z: BandedMatrix | None = None) -> BandedMatrix:
    """``Z .= alpha .* X .+ beta .* Y`` (src/generic/broadcast.jl:359-384, 927-964).  Without ``Z`` the result is allocated
    with the reference's broadcast bandwidths ``max.(bandwidths(X), bandwidths(Y))``; a given ``Z`` with fewer bands raises
    ``BandError`` when a dropped band holds a non-zero (checked on the device)."""
    if X.shape != Y.shape:
        raise DimensionMismatch(f"arrays could not be broadcast to a common size: {X.shape} and {Y.shape}")
    bw = (max(X.l, Y.l), max(X.u, Y.u))
    if Z is None:
        Z = BandedMatrix.undef(X.shape, bw, device=X.data.device)
    elif Z.shape != X.shape:
        raise DimensionMismatch(f"destination has size {Z.shape}, operands {X.shape}")
    elif Z.l < bw[0] or Z.u < bw[1]:
        for M_, c_ in ((X, alpha), (Y, beta)):  # non-zeros in bands the destination does not store -> BandError
            if c_ != 0 and (M_.l > Z.l or M_.u > Z.u):
                out = C.c_int64(0)
                hd = _h(Z.data)
                tmp = BandedMatrix.undef(Z.shape, (Z.l, Z.u), device=Z.data.device)
                hd.check(hd.lib.bmb200_dband_copy(hd.h, M_.m, M_.n, M_.l, M_.u, vp(M_.ptr), _ld_band(M_), tmp.l, tmp.u, vp(tmp.ptr),
                                                  _ld_band(tmp), C.byref(out)), "dband_copy")
                if out.value:
                    raise BandError(Z, (M_.l if M_.l > Z.l else -M_.u))
    if Z.m == 0 or Z.n == 0 or Z.l + Z.u + 1 <= 0:
        return Z
    hd = _h(Z.data)
    hd.check(hd.lib.bmb200_dband_axpby(hd.h, X.m, X.n, float(alpha), X.l, X.u, vp(X.ptr), _ld_band(X), float(beta), Y.l, Y.u,
                                       vp(Y.ptr), _ld_band(Y), Z.l, Z.u, vp(Z.ptr), _ld_band(Z)), "dband_axpby")
    return Z


def badd(A: BandedMatrix, B: BandedMatrix) -> BandedMatrix:
    """``A .+ B`` / ``A + B``."""
    return axpby_(1.0, A, 1.0, B)


def bsub(A: BandedMatrix, B: BandedMatrix) -> BandedMatrix:
    """``A .- B`` / ``A - B``."""
    return axpby_(1.0, A, -1.0, B)


def bscale(alpha, A: BandedMatrix) -> BandedMatrix:
    """``alpha .* A`` / ``alpha * A`` (same bandwidths)."""
    return axpby_(alpha, A, 0.0, A, similar(A))


def factorize(A: BandedMatrix):
    """_factorize (linalg.jl:75): square -> lu; rectangular -> qr (out of scope here)."""
    if A.m != A.n:
        raise NotImplementedError("rectangular banded solves use banded QR in the reference (out of scope)")
    return lu(A)


def solve(A, b: torch.Tensor) -> torch.Tensor:
    """``A \\ b`` (linalg.jl:5-9): copies b, checks squareness, factorises, ldiv!."""
    if isinstance(A, BandedCholesky):
        return ldiv_chol_(A, b.clone() if b.dim() == 1 else _clone_cm(b))
    if isinstance(A, (BandedLU, TransposeFact)):
        return ldiv_(A, b.clone() if b.dim() == 1 else _clone_cm(b))
    if A.m != A.n:  # checksquare
        raise DimensionMismatch(f"matrix is not square: dimensions are {A.shape}")
    if b.shape[0] != A.m:
        raise DimensionMismatch(f"B has first dimension {b.shape[0]} but needs {A.m}")
    return ldiv_(factorize(A), b.clone() if b.dim() == 1 else _clone_cm(b))


def _clone_cm(X: torch.Tensor) -> torch.Tensor:
    out = colmajor(X.shape[0], X.shape[1], device=X.device)
    out.copy_(X)
    return out


# ---------------------------------------------------------------------------------------------
# host-buffer entry points (the Fortran-ABI drop-in with HOST arrays; bench.py's e2e)
# ---------------------------------------------------------------------------------------------
def gbmv_host(trans, m, kl, ku, alpha, data: np.ndarray, x: np.ndarray, beta, y: np.ndarray, device=0):
    """BLAS.gbmv!(trans, m, kl, ku, α, data, x, β, y) on HOST arrays; data is (lda x n) Fortran-ordered."""
    hd = _lib.handle(device)
    n = data.shape[1]
    lda = data.strides[1] // 8 if n > 1 else max(1, data.shape[0])
    incx = x.strides[0] // 8 if x.size > 1 else 1
    incy = y.strides[0] // 8 if y.size > 1 else 1
    hd.check(hd.lib.bmb200_dgbmv_host(hd.h, trans.encode(), m, n, kl, ku, float(alpha), vp(data.ctypes.data), lda,
                                      vp(x.ctypes.data), incx, float(beta), vp(y.ctypes.data), incy), "dgbmv_host")
    return y
