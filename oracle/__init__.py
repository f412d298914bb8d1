"""CPU ORACLE for the banded hot path -- TEST INFRASTRUCTURE, not product code.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker.  The
product (``bandedmatrices.jl_b200``) never imports it and has no CPU fallback.

Two arithmetic back-ends, same call signatures:

* ``C``  -- ``libbmoracle.so`` built from ``oracle/bmoracle.c``: a plain-C restatement of the
  algorithms the reference delegates to (DGBMV, DGBTF2/DGBTRF, DGBTRS) plus the
  reference's own Julia driver loops (``_gbmm!``, widening copy).
* ``OB`` -- the OpenBLAS 0.3.30 ILP64 library that ships inside numpy in this image,
  entered through the very Fortran symbols BandedMatrices.jl ``ccall``s
  (``dgbmv_``: src/blas.jl:19-26, ``dgbtrf_``: src/banded/BandedLU.jl:98,
  ``dgbtrs_``: src/banded/linalg.jl:28,46,62).  Julia is not installed here, so this
  is the closest runnable piece of "the reference itself"; it PINS the C restatement
  (tests/test_oracle_pin.py) and is the CPU baseline timed by bench.py.

Above both sit restatements of the reference's driver logic (negative-bandwidth
re-viewing, ``gbmm!`` pruning, ``lu`` widening, ``ldiv!``), each citing file:line.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
i64 = C.c_int64
_dp = np.ctypeslib.ndpointer(dtype=np.float64)


def build(force: bool = False) -> None:
    """Compile libbmoracle.so / libblasdriver.so with the recipe in oracle/Makefile."""
    targets = [os.path.join(_HERE, "libbmoracle.so"), os.path.join(_HERE, "libblasdriver.so")]
    srcs = [os.path.join(_HERE, "bmoracle.c"), os.path.join(_HERE, "blasdriver.c")]
    stale = force or any(
        (not os.path.exists(t)) or os.path.getmtime(t) < os.path.getmtime(s) for t, s in zip(targets, srcs)
    )
    if stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])


def openblas_path() -> str:
    import numpy as _np

    hits = glob.glob(os.path.join(os.path.dirname(_np.__file__), "..", "numpy.libs", "libscipy_openblas64_*.so"))
    if not hits:
        raise RuntimeError("numpy's bundled ILP64 OpenBLAS not found")
    return os.path.abspath(hits[0])


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class _CBackend:
    name = "C"

    def __init__(self):
        build()
        L = C.CDLL(os.path.join(_HERE, "libbmoracle.so"))
        L.oracle_dgbmv.restype = C.c_int
        L.oracle_dgbmv.argtypes = [C.c_char, i64, i64, i64, i64, C.c_double, C.c_void_p, i64, C.c_void_p, i64,
                                   C.c_double, C.c_void_p, i64]
        L.oracle_gbmm.restype = C.c_int
        L.oracle_gbmm.argtypes = [i64] * 9 + [C.c_double, C.c_void_p, i64, C.c_void_p, i64, C.c_double, C.c_void_p, i64]
        L.oracle_dgbtrf.restype = C.c_int
        L.oracle_dgbtrf.argtypes = [i64, i64, i64, i64, C.c_void_p, i64, C.c_void_p]
        L.oracle_dgbtrs.restype = C.c_int
        L.oracle_dgbtrs.argtypes = [C.c_char, i64, i64, i64, i64, C.c_void_p, i64, C.c_void_p, C.c_void_p, i64]
        L.oracle_band_widen.restype = None
        L.oracle_band_widen.argtypes = [i64, i64, i64, C.c_void_p, i64, C.c_void_p, i64]
        L.oracle_banded_mul.restype = None
        L.oracle_banded_mul.argtypes = [i64] * 9 + [C.c_void_p, i64, C.c_void_p, i64, C.c_void_p, i64]
        L.oracle_fill_lmul.restype = None
        L.oracle_fill_lmul.argtypes = [C.c_double, C.c_void_p, i64, i64, i64]
        for f in (L.oracle_dtbsv, L.oracle_dtbmv):
            f.restype = C.c_int
            f.argtypes = [C.c_char, C.c_char, C.c_char, i64, i64, C.c_void_p, i64, C.c_void_p]
        L.oracle_dsbmv.restype = C.c_int
        L.oracle_dsbmv.argtypes = [C.c_char, i64, i64, C.c_double, C.c_void_p, i64, C.c_void_p, C.c_double, C.c_void_p]
        L.oracle_dpbtf2.restype = C.c_int
        L.oracle_dpbtf2.argtypes = [C.c_char, i64, i64, C.c_void_p, i64]
        L.oracle_dpbtrs.restype = C.c_int
        L.oracle_dpbtrs.argtypes = [C.c_char, i64, i64, i64, C.c_void_p, i64, C.c_void_p, i64]
        self.L = L

    def pbtrf(self, uplo, n, kd, ab, ldab):
        return self.L.oracle_dpbtf2(uplo.encode(), n, kd, _ptr(ab), ldab)

    def pbtrs(self, uplo, n, kd, nrhs, ab, ldab, b, ldb):
        return self.L.oracle_dpbtrs(uplo.encode(), n, kd, nrhs, _ptr(ab), ldab, _ptr(b), ldb)

    def sbmv(self, uplo, n, k, alpha, a, lda, x, beta, y):
        return self.L.oracle_dsbmv(uplo.encode(), n, k, alpha, _ptr(a), lda, _ptr(x), beta, _ptr(y))

    def tbsv(self, uplo, trans, diag, n, k, a, lda, x):
        return self.L.oracle_dtbsv(uplo.encode(), trans.encode(), diag.encode(), n, k, _ptr(a), lda, _ptr(x))

    def tbmv(self, uplo, trans, diag, n, k, a, lda, x):
        return self.L.oracle_dtbmv(uplo.encode(), trans.encode(), diag.encode(), n, k, _ptr(a), lda, _ptr(x))

    def gbmv_ptr(self, trans, m, n, kl, ku, alpha, a_addr, lda, x_addr, incx, beta, y_addr, incy):
        return self.L.oracle_dgbmv(trans.encode(), m, n, kl, ku, alpha, a_addr, lda, x_addr, incx, beta, y_addr, incy)

    def gbtrf(self, m, n, kl, ku, ab, ldab, ipiv):
        return self.L.oracle_dgbtrf(m, n, kl, ku, _ptr(ab), ldab, _ptr(ipiv))

    def gbtrs(self, trans, n, kl, ku, nrhs, ab, ldab, ipiv, b, ldb):
        return self.L.oracle_dgbtrs(trans.encode(), n, kl, ku, nrhs, _ptr(ab), ldab, _ptr(ipiv), _ptr(b), ldb)


class _OpenBLASBackend:
    """Fortran-ABI calls exactly as Julia issues them (by-reference Int64, hidden char lengths)."""

    name = "OpenBLAS"

    def __init__(self):
        self.L = C.CDLL(openblas_path())
        self.L.scipy_openblas_get_config64_.restype = C.c_char_p
        self.config = self.L.scipy_openblas_get_config64_().decode()
        self.set_threads(1)

    def set_threads(self, k: int) -> None:
        self.L.scipy_openblas_set_num_threads64_(C.c_int(int(k)))
        self.threads = int(k)

    def gbmv_ptr(self, trans, m, n, kl, ku, alpha, a_addr, lda, x_addr, incx, beta, y_addr, incy):
        if self.threads != 1 and kl + ku >= 15:
            raise RuntimeError("OpenBLAS threaded dgbmv overflows its work buffer for wide bands; pin 1 thread")
        r = C.byref
        self.L.scipy_dgbmv_64_(C.c_char_p(trans.encode()), r(i64(m)), r(i64(n)), r(i64(kl)), r(i64(ku)),
                               r(C.c_double(alpha)), C.c_void_p(a_addr), r(i64(lda)), C.c_void_p(x_addr),
                               r(i64(incx)), r(C.c_double(beta)), C.c_void_p(y_addr), r(i64(incy)), C.c_long(1))
        return 0

    def gbtrf(self, m, n, kl, ku, ab, ldab, ipiv):
        r = C.byref
        info = i64(0)
        self.L.scipy_dgbtrf_64_(r(i64(m)), r(i64(n)), r(i64(kl)), r(i64(ku)), _ptr(ab), r(i64(ldab)), _ptr(ipiv), r(info))
        return int(info.value)

    def _tb(self, fn, uplo, trans, diag, n, k, a, lda, x):
        r = C.byref
        fn(C.c_char_p(uplo.encode()), C.c_char_p(trans.encode()), C.c_char_p(diag.encode()), r(i64(n)), r(i64(k)), _ptr(a),
           r(i64(lda)), _ptr(x), r(i64(1)), C.c_long(1), C.c_long(1), C.c_long(1))
        return 0

    def sbmv(self, uplo, n, k, alpha, a, lda, x, beta, y):  # dsbmv_ as src/blas.jl:51-60 calls it
        r = C.byref
        self.L.scipy_dsbmv_64_(C.c_char_p(uplo.encode()), r(i64(n)), r(i64(k)), r(C.c_double(alpha)), _ptr(a), r(i64(lda)), _ptr(x),
                               r(i64(1)), r(C.c_double(beta)), _ptr(y), r(i64(1)), C.c_long(1))
        return 0

    def tbsv(self, uplo, trans, diag, n, k, a, lda, x):  # dtbsv_ as src/blas.jl:132-137 calls it
        return self._tb(self.L.scipy_dtbsv_64_, uplo, trans, diag, n, k, a, lda, x)

    def tbmv(self, uplo, trans, diag, n, k, a, lda, x):  # dtbmv_ as src/blas.jl:94-99 calls it
        return self._tb(self.L.scipy_dtbmv_64_, uplo, trans, diag, n, k, a, lda, x)

    def pbtrf(self, uplo, n, kd, ab, ldab):  # dpbtrf_ as pbtrf! calls it (src/lapack.jl:280-290)
        r = C.byref
        info = i64(0)
        self.L.scipy_dpbtrf_64_(C.c_char_p(uplo.encode()), r(i64(n)), r(i64(kd)), _ptr(ab), r(i64(ldab)), r(info), C.c_long(1))
        return int(info.value)

    def pbtrs(self, uplo, n, kd, nrhs, ab, ldab, b, ldb):  # dpbtrs_ as pbtrs! calls it (src/lapack.jl:312-326)
        r = C.byref
        info = i64(0)
        self.L.scipy_dpbtrs_64_(C.c_char_p(uplo.encode()), r(i64(n)), r(i64(kd)), r(i64(nrhs)), _ptr(ab), r(i64(ldab)), _ptr(b),
                                r(i64(ldb)), r(info), C.c_long(1))
        return int(info.value)

    def gbtf2(self, m, n, kl, ku, ab, ldab, ipiv):
        r = C.byref
        info = i64(0)
        self.L.scipy_dgbtf2_64_(r(i64(m)), r(i64(n)), r(i64(kl)), r(i64(ku)), _ptr(ab), r(i64(ldab)), _ptr(ipiv), r(info))
        return int(info.value)

    def gbtrs(self, trans, n, kl, ku, nrhs, ab, ldab, ipiv, b, ldb):
        r = C.byref
        info = i64(0)
        self.L.scipy_dgbtrs_64_(C.c_char_p(trans.encode()), r(i64(n)), r(i64(kl)), r(i64(ku)), r(i64(nrhs)), _ptr(ab),
                                r(i64(ldab)), _ptr(ipiv), _ptr(b), r(i64(ldb)), r(info), C.c_long(1))
        return int(info.value)

    # ---- S / C / Z instantiations (src/blas.jl:4-7; LAPACK.gbtrf! / gbtrs! for the four element types): the OpenBLAS entry points
    # themselves, selected by the numpy dtype of the arrays.  alpha / beta go by reference as one element of the type.
    @staticmethod
    def prefix(dtype):
        return {"float32": "s", "float64": "d", "complex64": "c", "complex128": "z"}[np.dtype(dtype).name]

    def t_gbmv(self, trans, m, n, kl, ku, alpha, a, lda, x, beta, y):
        p, r = self.prefix(a.dtype), C.byref
        al, be = np.array([alpha], dtype=a.dtype), np.array([beta], dtype=a.dtype)
        getattr(self.L, f"scipy_{p}gbmv_64_")(C.c_char_p(trans.encode()), r(i64(m)), r(i64(n)), r(i64(kl)), r(i64(ku)), _ptr(al), _ptr(a),
                                             r(i64(lda)), _ptr(x), r(i64(1)), _ptr(be), _ptr(y), r(i64(1)), C.c_long(1))
        return 0

    def t_hbmv(self, uplo, n, k, alpha, a, lda, x, beta, y):
        p, r = self.prefix(a.dtype), C.byref
        name = f"scipy_{p}sbmv_64_" if p in "sd" else f"scipy_{p}hbmv_64_"
        al, be = np.array([alpha], dtype=a.dtype), np.array([beta], dtype=a.dtype)
        getattr(self.L, name)(C.c_char_p(uplo.encode()), r(i64(n)), r(i64(k)), _ptr(al), _ptr(a), r(i64(lda)), _ptr(x), r(i64(1)), _ptr(be),
                              _ptr(y), r(i64(1)), C.c_long(1))
        return 0

    def t_gbtrf(self, m, n, kl, ku, ab, ldab, ipiv):
        p, r = self.prefix(ab.dtype), C.byref
        info = i64(0)
        getattr(self.L, f"scipy_{p}gbtrf_64_")(r(i64(m)), r(i64(n)), r(i64(kl)), r(i64(ku)), _ptr(ab), r(i64(ldab)), _ptr(ipiv), r(info))
        return int(info.value)

    def t_gbtrs(self, trans, n, kl, ku, nrhs, ab, ldab, ipiv, b, ldb):
        p, r = self.prefix(ab.dtype), C.byref
        info = i64(0)
        getattr(self.L, f"scipy_{p}gbtrs_64_")(C.c_char_p(trans.encode()), r(i64(n)), r(i64(kl)), r(i64(ku)), r(i64(nrhs)), _ptr(ab),
                                              r(i64(ldab)), _ptr(ipiv), _ptr(b), r(i64(ldb)), r(info), C.c_long(1))
        return int(info.value)
    def t_tb(self, name, uplo, trans, diag, n, k, a, lda, x):  # {s,c,z}tbsv_ / tbmv_ as src/blas.jl:94-99, 132-137 call them
        p, r = self.prefix(a.dtype), C.byref
        getattr(self.L, f"scipy_{p}{name}_64_")(C.c_char_p(uplo.encode()), C.c_char_p(trans.encode()), C.c_char_p(diag.encode()), r(i64(n)), r(i64(k)),
                                               _ptr(a), r(i64(lda)), _ptr(x), r(i64(1)), C.c_long(1), C.c_long(1), C.c_long(1))
        return 0

    def t_pbtrf(self, uplo, n, kd, ab, ldab):
        p, r = self.prefix(ab.dtype), C.byref
        info = i64(0)
        getattr(self.L, f"scipy_{p}pbtrf_64_")(C.c_char_p(uplo.encode()), r(i64(n)), r(i64(kd)), _ptr(ab), r(i64(ldab)), r(info), C.c_long(1))
        return int(info.value)

    def t_pbtrs(self, uplo, n, kd, nrhs, ab, ldab, b, ldb):
        p, r = self.prefix(ab.dtype), C.byref
        info = i64(0)
        getattr(self.L, f"scipy_{p}pbtrs_64_")(C.c_char_p(uplo.encode()), r(i64(n)), r(i64(kd)), r(i64(nrhs)), _ptr(ab), r(i64(ldab)), _ptr(b),
                                              r(i64(ldb)), r(info), C.c_long(1))
        return int(info.value)


_backends: dict = {}


def backend(name: str = "C"):
    if name not in _backends:
        _backends[name] = _CBackend() if name == "C" else _OpenBLASBackend()
    return _backends[name]


# --------------------------------------------------------------------------------------
# Host band container used by the oracle's driver restatements.
# --------------------------------------------------------------------------------------
@dataclass
class Band:
    """data[(u+k-j), j] = A[k,j] (0-based); src/banded/BandedMatrix.jl:16-28, :414-419."""

    data: np.ndarray  # (l+u+1) x n, Fortran order (may be a column-strided view: lda = data.strides[1]//8)
    m: int
    l: int
    u: int

    @property
    def n(self):
        return self.data.shape[1]

    @property
    def lda(self):
        return self.data.strides[1] // 8 if self.data.shape[1] > 1 else max(1, self.data.shape[0])

    def dense(self) -> np.ndarray:
        out = np.zeros((self.m, self.n))
        for j in range(self.n):
            for k in range(max(0, j - self.u), min(self.m - 1, j + self.l) + 1):
                out[k, j] = self.data[self.u + k - j, j]
        return out


def band_from_dense(A: np.ndarray, l: int, u: int, fill=np.nan) -> Band:
    """BandedMatrix(A,(l,u)) (BandedMatrix.jl:222-232); out-of-matrix corner slots get ``fill``."""
    m, n = A.shape
    data = np.full((max(0, l + u + 1), n), fill, order="F")
    for j in range(n):
        for k in range(max(0, j - u), min(m - 1, j + l) + 1):
            data[u + k - j, j] = A[k, j]
    return Band(data, m, l, u)


def brand(rng, m, n, l, u, corners=None) -> Band:
    """brand(m,n,l,u): uniform over the WHOLE data array (BandedMatrix.jl:184-187).
    ``corners`` (e.g. NaN) overwrites the out-of-matrix slots to prove they are never read."""
    data = np.asfortranarray(rng.random((max(0, l + u + 1), n)))
    if corners is not None:
        for j in range(n):
            for r in range(data.shape[0]):
                k = j + r - u
                if k < 0 or k >= m:
                    data[r, j] = corners
    return Band(data, m, l, u)


def _addr(a: np.ndarray, off_elems: int = 0) -> int:
    return a.ctypes.data + 8 * off_elems


def gbmv(be, trans, m, kl, ku, alpha, data, x, beta, y):
    """BLAS.gbmv!(trans, m, kl, ku, alpha, data, x, beta, y): n = size(data,2), lda = stride(data,2)
    (stdlib wrapper reached from src/generic/matmul.jl:21-23).  x,y may be strided 1-D views."""
    n = data.shape[1]
    lda = data.strides[1] // 8 if n > 1 else max(1, data.shape[0])
    incx = x.strides[0] // 8 if x.size > 1 else 1
    incy = y.strides[0] // 8 if y.size > 1 else 1
    be.gbmv_ptr(trans, m, n, kl, ku, alpha, _addr(data), lda, _addr(x), incx, beta, _addr(y), incy)
    return y


def _fill_rmul(y, beta):
    """_fill_rmul!(y, beta): src/generic/utils.jl:30."""
    if beta == 0:
        y[...] = 0.0
    else:
        y *= beta
    return y


def banded_muladd_vec(be, alpha, A: Band, x, beta, y):
    """_banded_muladd!(alpha,A,x,beta,y) + _banded_gbmv! : src/generic/matmul.jl:26-59."""
    m, n, l, u = A.m, A.n, A.l, A.u
    if x.shape[0] != n or y.shape[0] != m:
        raise ValueError("DimensionMismatch")

    def _gbmv(mm, kl, ku, data, xv, yv):  # _banded_gbmv! :26-39
        if yv.shape[0] == 0:
            return
        if xv.shape[0] == 0:
            _fill_rmul(yv, beta)
            return
        gbmv(be, "N", mm, kl, ku, alpha, data, np.array(xv, copy=True), beta, yv)

    if -l > u:
        _fill_rmul(y, beta)
    elif l < 0:  # view(A,:,1-l:n): bandwidths (0, u+l); data rows 0..u+l of columns -l..n-1
        _gbmv(m, 0, u + l, A.data[: u + l + 1, -l:], x[-l:], y)
    elif u < 0:  # view(A,1-u:m,:): bandwidths (l+u, 0); same data rows (row index = k'-j with u'=0)
        _fill_rmul(y[:-u], beta)
        _gbmv(m + u, l + u, 0, A.data, x, y[-u:])
    else:
        _gbmv(m, l, u, A.data, x, y)
    return y


def gbmm_kernel(be, alpha, A_data, B_data, beta, C_data, n, nu, m, Al, Au, Bl, Bu, Cl, Cu):
    """_gbmm! (src/banded/gbmm.jl:296-340) replayed call-for-call on backend ``be``."""
    sta = A_data.strides[1] // 8 if A_data.shape[1] > 1 else A_data.shape[0]
    stb = B_data.strides[1] // 8 if B_data.shape[1] > 1 else B_data.shape[0]
    stc = C_data.strides[1] // 8 if C_data.shape[1] > 1 else C_data.shape[0]
    a, b, c = _addr(A_data), _addr(B_data), _addr(C_data)
    g = be.gbmv_ptr
    for j in range(1, min(m, 1 + Bu) + 1):
        g("N", min(Cl + j, n), min(Bl + j, nu), Al, Au, alpha, a, sta, b + 8 * ((j - 1) * stb + Bu - j + 1), 1, beta,
          c + 8 * ((j - 1) * stc + Cu - j + 1), 1)
    for j in range(2 + Bu, min(1 + Cu, nu + Bu, m) + 1):
        g("N", min(Cl + j, n), min(Bl + Bu + 1, nu - j + Bu + 1), Al + j - Bu - 1, Au - j + Bu + 1, alpha,
          a + 8 * (j - Bu - 1) * sta, sta, b + 8 * (j - 1) * stb, 1, beta, c + 8 * ((j - 1) * stc + Cu - j + 1), 1)
    for j in range(2 + Cu, min(m, nu + Bu, n + Cu) + 1):
        p = j - Bu
        g("N", min(Cl + Cu + 1, n - j + Cu + 1), min(Bl + Bu + 1, nu - p + 1), Al + Au, 0, alpha,
          a + 8 * (j - Bu - 1) * sta, sta, b + 8 * (j - 1) * stb, 1, beta, c + 8 * (j - 1) * stc, 1)
    j0, j1 = nu + Bu + 1, min(m, n + Cu)
    if j1 >= j0:
        _fill_rmul(C_data[:, j0 - 1 : j1], beta)
    return C_data


def lu(be, A: Band):
    """lu(A): widen to (l, l+u) storage then gbtrf! (src/banded/BandedLU.jl:90-111).
    Returns (AB, ipiv[1-based int64], info)."""
    if A.m != A.n:
        raise ValueError("DimensionMismatch: matrix is not square")
    n, l, u = A.n, A.l, A.u
    ab = np.zeros((2 * l + u + 1, n), order="F")
    ab[l:, :] = A.data
    ipiv = np.zeros(n, dtype=np.int64)
    info = be.gbtrf(n, n, l, u, ab, ab.shape[0], ipiv) if n > 0 else 0
    return ab, ipiv, info


def ldiv(be, trans, ab, ipiv, l, u, B):
    """ldiv!(F, B) -> gbtrs!(trans, l, u, m, data, ipiv, B)  (src/banded/linalg.jl:24-30, :41-47)."""
    n = ab.shape[1]
    Bm = B.reshape(n, -1) if B.ndim == 1 else B
    assert Bm.flags.f_contiguous or Bm.shape[1] == 1
    if n > 0:
        be.gbtrs(trans, n, l, u, Bm.shape[1], ab, ab.shape[0], ipiv, Bm, max(1, n))
    return B
