"""GPU parity of the band-aligned elementwise operations between different bandwidths:
banded_axpy! (src/banded/BandedMatrix.jl:1006-1015; src/generic/broadcast.jl:978-1020) and copyto! (broadcast.jl:175-230).
The reference does these in Julia itself, so the check is its own arithmetic restated in numpy: unequal bandwidths
``a*x + y`` (a rounded product, then a sum) -- bit-identical; equal bandwidths ``axpy!(a, X.data, Y.data)`` = OpenBLAS daxpy,
one FMA per slot -- compared with OpenBLAS' daxpy_ itself, bit for bit."""
import ctypes as C
import itertools

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def _dense(data, m, l, u):
    return oracle.Band(np.asfortranarray(data), m, l, u).dense()


@pytest.mark.parametrize("shape", [(50, 50), (300, 280), (280, 300), (4000, 4000)])
@pytest.mark.parametrize("bands", [((2, 3), (4, 5)), ((1, 1), (1, 4)), ((0, 2), (3, 2)), ((3, 0), (3, 0)), ((-1, 2), (1, 3)), ((40, 33), (64, 64))])
def test_axpy_and_copyto_between_bandwidths(bm, rng, shape, bands):
    m, n = shape
    (xl, xu), (yl, yu) = bands
    X = oracle.brand(rng, m, n, xl, xu, corners=np.nan)
    Y = oracle.brand(rng, m, n, yl, yu, corners=7.5)
    a = 0.37
    dX, dY = bm.BandedMatrix.from_banddata(X.data, m, xl, xu), bm.BandedMatrix.from_banddata(Y.data, m, yl, yu)
    if (xl, xu) == (yl, yu):
        ref = Y.data.copy(order="F")
        xd = np.where(np.isnan(X.data), 1.25, X.data)  # daxpy touches the corner slots too: give them values
        dX = bm.BandedMatrix.from_banddata(xd, m, xl, xu)
        ob = oracle.backend("OB")
        r = C.byref
        ob.L.scipy_daxpy_64_(r(C.c_int64(ref.size)), r(C.c_double(a)), np.asfortranarray(xd).ctypes.data_as(C.c_void_p), r(C.c_int64(1)),
                             ref.ctypes.data_as(C.c_void_p), r(C.c_int64(1)))
        bm.axpy_(a, dX, dY)
        assert np.array_equal(dY.banddata_host(), ref)
    else:
        refd = a * X.dense() + Y.dense()           # numpy: product rounded, then sum -- the reference's scalar loop
        bm.axpy_(a, dX, dY)
        got = dY.banddata_host()
        assert np.array_equal(_dense(got, m, yl, yu), refd)
        mask = np.isnan(oracle.brand(rng, m, n, yl, yu, corners=np.nan).data)
        assert np.all(got[mask] == 7.5)             # corner slots of Y untouched
    # copyto!: overlapping bands copied, the rest of dest's bands zeroed
    dD = bm.BandedMatrix.from_banddata(np.full((yl + yu + 1, n), 3.0), m, yl, yu)
    bm.copyto_(dD, bm.BandedMatrix.from_banddata(X.data, m, xl, xu))
    assert np.array_equal(_dense(dD.banddata_host(), m, yl, yu), X.dense())


def test_band_error_when_nonzeros_fall_outside(bm, rng):
    m = n = 200
    X = oracle.brand(rng, m, n, 3, 2)
    dX = bm.BandedMatrix.from_banddata(X.data, m, 3, 2)
    Y0 = oracle.brand(rng, m, n, 2, 2)
    dY = bm.BandedMatrix.from_banddata(Y0.data, m, 2, 2)
    with pytest.raises(bm.BandError):
        bm.axpy_(2.0, dX, dY)
    assert np.array_equal(dY.banddata_host(), Y0.data)      # nothing written before the error (broadcast.jl:991-1006)
    with pytest.raises(bm.BandError):
        bm.copyto_(dY, dX)
    X.data[-1, :] = 0.0                                      # the third sub-diagonal is structurally zero: now legal
    dX = bm.BandedMatrix.from_banddata(X.data, m, 3, 2)
    bm.axpy_(2.0, dX, dY)
    assert np.array_equal(_dense(dY.banddata_host(), m, 2, 2), 2.0 * X.dense() + Y0.dense())
    with pytest.raises(bm.DimensionMismatch):
        bm.axpy_(1.0, dX, bm.BandedMatrix.zeros((m, n + 1), (3, 2)))


def _rand_banded_dense(rng, m, n, l, u):
    D = rng.standard_normal((m, n))
    k, j = np.meshgrid(np.arange(m), np.arange(n), indexing="ij")
    D[(k - j > l) | (j - k > u)] = 0.0
    return D


def test_materialize_transpose_and_zeroband_counts(bm, rng):
    """convert(BandedMatrix, A') and gbmm.jl:191-205's zero-band counts, computed by device kernels."""
    from bandedmatrices_b200.linalg import _num_zeroband_l, _num_zeroband_u, materialize_transpose

    for (m, n, l, u) in [(8, 11, 2, 3), (11, 8, 0, 4), (6, 6, 1, 0), (300, 200, 17, 40), (1, 1, 0, 0)]:
        D = _rand_banded_dense(rng, m, n, l, u)
        A = bm.BandedMatrix.from_dense(D, (l, u))
        T = materialize_transpose(A.T)
        assert (T.l, T.u) == (u, l)
        assert np.array_equal(T.to_dense(), D.T)
    D = _rand_banded_dense(rng, 9, 9, 2, 3)
    D[np.arange(6), np.arange(6) + 3] = 0  # top band all zero
    D[np.arange(7), np.arange(7) + 2] = 0
    A = bm.BandedMatrix.from_dense(D, (2, 3))
    assert _num_zeroband_u(A) == 2 and _num_zeroband_l(A) == 0
    Z = bm.BandedMatrix.from_dense(np.zeros((5, 5)), (1, 1))
    assert _num_zeroband_u(Z) == 3 and _num_zeroband_l(Z) == 3


@pytest.mark.parametrize("bands", [((2, 3), (2, 3)), ((1, 4), (3, 0)), ((0, 0), (5, 2)), ((7, 1), (-1, 3)), ((40, 33), (2, 2))])
def test_broadcast_axpby_add_sub_scale(bm, rng, bands):
    """A .+ B, A .- B, a .* A, a .* A .+ b .* B (src/generic/broadcast.jl:359-384, 927-964): result bandwidths are the
    element-wise maxima, every entry is round(round(a*x) + round(b*y)) -- numpy's arithmetic on the dense forms."""
    (xl, xu), (yl, yu) = bands
    m, n = 120, 97
    X = _rand_banded_dense(rng, m, n, xl, xu)
    Y = _rand_banded_dense(rng, m, n, yl, yu)
    A, B = bm.BandedMatrix.from_dense(X, (xl, xu)), bm.BandedMatrix.from_dense(Y, (yl, yu))
    S = bm.badd(A, B)
    assert (S.l, S.u) == (max(xl, yl), max(xu, yu))
    assert np.array_equal(S.to_dense(), X + Y)
    assert np.array_equal(bm.bsub(A, B).to_dense(), X - Y)
    assert np.array_equal(bm.bscale(-2.5, A).to_dense(), -2.5 * X)
    assert np.array_equal(bm.axpby_(3.0, A, 0.125, B).to_dense(), 3.0 * X + 0.125 * Y)
    # a destination with more bands gets zeros there; with fewer bands a dropped non-zero is a BandError
    Zw = bm.BandedMatrix.from_dense(rng.standard_normal((m, n)), (max(xl, yl) + 2, max(xu, yu) + 1))
    bm.axpby_(1.0, A, 1.0, B, Zw)
    assert np.array_equal(Zw.to_dense(), X + Y)
    if max(xl, yl) >= 1:
        Zn = bm.BandedMatrix.zeros((m, n), (max(xl, yl) - 1, max(xu, yu)))
        with pytest.raises(bm.BandError):
            bm.axpby_(1.0, A, 1.0, B, Zn)
    assert bm.similar(A).shape == A.shape and (bm.similar(A, (1, 1)).l, bm.similar(A, (1, 1)).u) == (1, 1)
