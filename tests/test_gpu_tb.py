"""GPU parity of the triangular band solve / multiply (tbsv! / tbmv!, src/blas.jl:71-141; ldiv! / lmul! of
UpperTriangular / LowerTriangular{<:BandedMatrix}, src/tribanded.jl:47-84) through the C ABI: bit-identical to the oracle
(itself pinned bit-for-bit to OpenBLAS dtbsv_ / dtbmv_)."""
import itertools

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _tri_band(rng, n, k, uplo, lda_extra=0):
    a = np.asfortranarray(rng.standard_normal((k + 1 + lda_extra, n))) / (2 * k + 2)
    a[k if uplo == "U" else 0, :] = (1.0 + rng.random(n)) * rng.choice([-1.0, 1.0], n)
    return a


def _dev(a):
    """(rows x n) Fortran band array -> the package's (n, rows) tensor with the same memory layout."""
    return torch.as_tensor(np.ascontiguousarray(a.T)).cuda()


@pytest.mark.parametrize("shape", [(1, 0), (5, 2), (40, 3), (1000, 16), (4097, 40), (30000, 7), (6000, 300), (50, 80),
                                   (20011, 1024), (3000, 1500), (2, 1), (700, 63), (700, 64), (129, 127)])
def test_tbsv_tbmv_bit_identical(bm, oracle_c, rng, shape):
    n, k = shape
    for uplo, diag, extra in itertools.product("UL", "NU", (0, 3)):
        a = _tri_band(rng, n, k, uplo, extra)
        lda = a.shape[0]
        dA = _dev(a)
        for name, fn in (("tbsv", bm.tbsv_), ("tbmv", bm.tbmv_)):
            x0 = rng.standard_normal(n)
            ref = x0.copy()
            assert getattr(oracle_c, name)(uplo, "N", diag, n, k, a, lda, ref) == 0
            x = torch.as_tensor(x0).cuda()
            fn(uplo, "N", diag, n, k, dA, x)
            assert np.array_equal(x.cpu().numpy(), ref), (name, uplo, diag, n, k, extra)


@pytest.mark.parametrize("shape", [(1, 0), (5, 2), (40, 3), (1000, 16), (4097, 40), (30000, 7), (3000, 300), (50, 80), (2500, 1024),
                                   (2, 1), (700, 63), (700, 64), (3000, 1500)])
def test_tbsv_tbmv_transposed(bm, oracle_c, rng, shape):
    """trans = 'T' (row-major layouts, src/tribanded.jl:86-96): dot-product forms; OpenBLAS' own summation order is
    unspecified there, so the comparison with the oracle is to 1e-13 (relative to max|x|), like the other 'T' paths."""
    n, k = shape
    for uplo, diag, extra in itertools.product("UL", "NU", (0, 3)):
        a = _tri_band(rng, n, k, uplo, extra)
        lda = a.shape[0]
        dA = _dev(a)
        for name, fn in (("tbsv", bm.tbsv_), ("tbmv", bm.tbmv_)):
            x0 = rng.standard_normal(n)
            ref = x0.copy()
            assert getattr(oracle_c, name)(uplo, "T", diag, n, k, a, lda, ref) == 0
            x = torch.as_tensor(x0).cuda()
            fn(uplo, "T", diag, n, k, dA, x)
            got = x.cpu().numpy()
            assert np.max(np.abs(got - ref)) <= 1e-13 * max(1.0, np.max(np.abs(ref))), (name, uplo, diag, n, k, extra)


def test_triangular_views_of_a_banded_matrix(bm, oracle_c, rng):
    """ldiv!(UpperTriangular(A), x) etc.: the triangular views share A's data array (rows 1:u+1 / u+1:u+l+1,
    src/tribanded.jl:47-84), lda = l+u+1."""
    n, l, u = 5000, 37, 21
    data = np.asfortranarray(rng.standard_normal((l + u + 1, n))) / (2 * (l + u))
    data[u, :] = 2.0 + rng.random(n)
    A = bm.BandedMatrix.from_banddata(data, n, l, u)
    for uplo, unit in itertools.product("UL", (False, True)):
        a = data[: u + 1, :] if uplo == "U" else data[u:, :]
        k = u if uplo == "U" else l
        for name, fn in (("tbsv", bm.ldiv_tri_), ("tbmv", bm.lmul_tri_)):
            x0 = rng.standard_normal(n)
            ref = x0.copy()
            # the oracle gets the same strided view: lda = l+u+1
            getattr(oracle_c, name)(uplo, "N", "U" if unit else "N", n, k, a, l + u + 1, ref)
            x = torch.as_tensor(x0).cuda()
            fn(uplo, unit, A, x)
            assert np.array_equal(x.cpu().numpy(), ref), (name, uplo, unit)


@pytest.mark.parametrize("shape", [(1, 0), (5, 2), (40, 3), (1000, 16), (4097, 40), (100000, 4), (6000, 300), (50, 80), (20011, 1024)])
def test_sbmv_matches_oracle(bm, oracle_c, rng, shape):
    """sbmv! (src/blas.jl:36-66) / mul!(y, Symmetric(A), x) (symbanded.jl:72-93): OpenBLAS' summation order is unspecified
    (axpy + SIMD dot per column), so 1e-13 relative to max|y|; beta == 0 overwrites NaN; alpha == 0 does not read A."""
    n, k = shape
    for uplo, (al, be), extra in itertools.product("UL", [(1.0, 0.0), (0.7, -1.3), (0.0, 2.0)], (0, 3)):
        a = np.asfortranarray(rng.standard_normal((k + 1 + extra, n)))
        lda = a.shape[0]
        if al == 0.0:
            a[:] = np.nan
        x0 = rng.standard_normal(n)
        y0 = rng.standard_normal(n)
        if be == 0.0:
            y0[:] = np.nan
        ref = y0.copy()
        assert oracle_c.sbmv(uplo, n, k, al, a, lda, x0, be, ref) == 0
        y = torch.as_tensor(y0).cuda()
        bm.sbmv_(uplo, k, al, _dev(a), torch.as_tensor(x0).cuda(), be, y)
        got = y.cpu().numpy()
        assert np.isfinite(got).all()
        assert np.max(np.abs(got - ref)) <= 1e-13 * max(1.0, np.max(np.abs(ref))), (uplo, al, be, n, k)


def test_mul_symmetric_view(bm, rng):
    """mul!(y, Symmetric(A, uplo), x): reads only the `uplo` triangle of A's data; x === y is un-aliased first."""
    n, l, u = 3000, 9, 14
    data = np.asfortranarray(rng.standard_normal((l + u + 1, n)))
    A = bm.BandedMatrix.from_banddata(data, n, l, u)
    D = A.to_dense()
    for uplo in "UL":
        T = np.triu(D) if uplo == "U" else np.tril(D)
        S = T + T.T - np.diag(np.diag(D))
        x0 = rng.standard_normal(n)
        y = torch.full((n,), float("nan"), dtype=torch.float64, device="cuda")
        bm.mul_sym_(y, uplo, A, torch.as_tensor(x0).cuda())
        assert np.allclose(y.cpu().numpy(), S @ x0, rtol=1e-12, atol=1e-12)
        z = torch.as_tensor(x0).cuda()
        bm.mul_sym_(z, uplo, A, z, 2.0, 0.0)  # aliased
        assert np.allclose(z.cpu().numpy(), 2.0 * (S @ x0), rtol=1e-12, atol=1e-12)


def test_tb_argument_errors(bm, rng):
    n, k = 10, 2
    dA = _dev(_tri_band(rng, n, k, "U"))
    x = torch.zeros(n, dtype=torch.float64, device="cuda")
    with pytest.raises(bm.DimensionMismatch):
        bm.tbsv_("U", "N", "N", n + 1, k, dA, x)
    with pytest.raises(bm.DimensionMismatch):
        bm.tbmv_("U", "N", "N", n, k, dA, x[:-1])
    with pytest.raises(ValueError):
        bm.tbsv_("U", "N", "N", n, k + 1, dA, x)
    with pytest.raises(bm.BMB200Error):  # invalid trans
        bm.tbsv_("U", "X", "N", n, k, dA, x)
    assert bm.tbsv_("L", "N", "U", 0, 0, torch.zeros((0, 1), dtype=torch.float64, device="cuda"), x[:0]).numel() == 0
