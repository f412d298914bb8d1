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
                                   (20011, 1024), (3000, 1500)])
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


@pytest.mark.parametrize("shape", [(1, 0), (5, 2), (40, 3), (1000, 16), (4097, 40), (30000, 7), (3000, 300), (50, 80), (2500, 1024)])
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
