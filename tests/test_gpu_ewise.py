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
