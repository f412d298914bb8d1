"""GPU parity of the banded Cholesky (pbtrf! / pbtrs!, src/lapack.jl:268-332; cholesky / ldiv! of Symmetric{<:BandedMatrix},
src/symbanded/BandedCholesky.jl) through the C ABI.  kd <= 64 (DPBTF2, the algorithm DPBTRF runs there): factors bit-identical
to the oracle, itself pinned bit-for-bit to OpenBLAS dpbtrf_ (tests/test_oracle_pin.py).  Wider bands (blocked; DPBTRF's
DSYRK/DGEMM order is unspecified): 1e-12 relative to OpenBLAS' factor and ||U'U - A|| <= 1e-14 kd ||A||."""
import itertools

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _spd_band(rng, n, kd, uplo, lda_extra=0):
    ab = np.asfortranarray(rng.standard_normal((kd + 1 + lda_extra, n)))
    ab[kd if uplo == "U" else 0, :] = 2.0 * (kd + 1) + rng.random(n)
    return ab


def _dev(a):
    return torch.as_tensor(np.ascontiguousarray(a.T)).cuda()


def _host(t):
    return np.asfortranarray(t.cpu().numpy().T)


def _sym_dense(ab, n, kd, uplo):
    A = np.zeros((n, n))
    for k in range(n):
        for d in range(0, min(kd, k if uplo == "U" else n - 1 - k) + 1):
            if uplo == "U":
                A[k - d, k] = A[k, k - d] = ab[kd - d, k]
            else:
                A[k + d, k] = A[k, k + d] = ab[d, k]
    return A


@pytest.mark.parametrize("shape", [(1, 0), (2, 1), (7, 2), (40, 3), (1000, 4), (3000, 8), (2000, 9), (5000, 16), (1500, 17),
                                   (4000, 31), (2500, 32), (1200, 33), (900, 64), (20, 40), (64, 63), (30000, 5)])
@pytest.mark.parametrize("diag_kernel", [0, -1, 1])
def test_pbtrf_pbtrs_narrow_bit_identical(bm, oracle_c, rng, shape, diag_kernel):
    """diag_kernel: tuning value of pb_nodiag -- 0 shipped dispatch, -1 the one-warp register kernel for every kd <= 31, 1 the
    shared-memory window kernel for every kd <= 64: both kernels must reproduce DPBTF2 bit for bit."""
    n, kd = shape
    bm.handle(0).tune("pb_nodiag", diag_kernel)
    try:
        _narrow_case(bm, oracle_c, rng, n, kd)
    finally:
        bm.handle(0).tune("reset", 0)


def _narrow_case(bm, oracle_c, rng, n, kd):
    for uplo, extra in itertools.product("UL", (0, 3)):
        ab = _spd_band(rng, n, kd, uplo, extra)
        ldab = ab.shape[0]
        ref = ab.copy(order="F")
        assert oracle_c.pbtrf(uplo, n, kd, ref, ldab) == 0
        dA = _dev(ab)
        _, info = bm.pbtrf_(uplo, n, kd, dA)
        assert info == 0
        got = _host(dA)
        assert np.array_equal(got[: kd + 1], ref[: kd + 1]), (uplo, n, kd, extra)
        if extra:
            assert np.array_equal(got[kd + 1:], ab[kd + 1:])  # rows beyond the band array are not touched
        nrhs = 3
        b = np.asfortranarray(rng.standard_normal((n, nrhs)))
        bref = b.copy(order="F")
        assert oracle_c.pbtrs(uplo, n, kd, nrhs, ref, ldab, bref, n) == 0
        dB = bm.to_colmajor(b)
        bm.pbtrs_(uplo, n, kd, dA, dB)
        gotb = dB.cpu().numpy()
        assert np.max(np.abs(gotb - bref)) <= 1e-13 * max(1.0, np.max(np.abs(bref))), (uplo, n, kd)
        x = torch.as_tensor(b[:, 0].copy()).cuda()
        bm.pbtrs_(uplo, n, kd, dA, x)
        assert np.array_equal(x.cpu().numpy(), gotb[:, 0])


@pytest.mark.parametrize("shape", [(300, 65), (66, 65), (1000, 100), (700, 128), (2000, 300), (3000, 1024), (500, 499), (130, 129), (4096, 64 * 3)])
def test_pbtrf_wide_blocked(bm, oracle_ob, rng, shape):
    n, kd = shape
    for uplo, extra in itertools.product("UL", (0, 1)):
        ab = _spd_band(rng, n, kd, uplo, extra)
        ldab = ab.shape[0]
        ref = ab.copy(order="F")
        assert oracle_ob.pbtrf(uplo, n, kd, ref, ldab) == 0
        dA = _dev(ab)
        _, info = bm.pbtrf_(uplo, n, kd, dA)
        assert info == 0
        got = _host(dA)
        # in-matrix entries only (the unused corner of the band array is never read or written)
        mask = np.zeros_like(ref[: kd + 1], dtype=bool)
        for d in range(kd + 1):
            if uplo == "U":
                mask[kd - d, d:] = True
            else:
                mask[d, : n - d] = True
        err = np.max(np.abs(got[: kd + 1][mask] - ref[: kd + 1][mask]))
        assert err <= 1e-12 * np.max(np.abs(ref[: kd + 1][mask])), (uplo, n, kd, err)
        assert np.array_equal(got[: kd + 1][~mask], ab[: kd + 1][~mask])
        if n <= 1000:
            A = _sym_dense(ab, n, kd, uplo)
            F = _sym_dense(got, n, kd, uplo)
            U = np.triu(F) if uplo == "U" else np.tril(F).T
            assert np.max(np.abs(U.T @ U - A)) <= 1e-14 * kd * np.max(np.abs(A))
        b = np.asfortranarray(rng.standard_normal((n, 2)))
        bref = b.copy(order="F")
        assert oracle_ob.pbtrs(uplo, n, kd, 2, ref, ldab, bref, n) == 0
        dB = bm.to_colmajor(b)
        bm.pbtrs_(uplo, n, kd, dA, dB)
        assert np.max(np.abs(dB.cpu().numpy() - bref)) <= 1e-12 * max(1.0, np.max(np.abs(bref))), (uplo, n, kd)


@pytest.mark.parametrize("shape", [(700, 128), (2000, 300), (3000, 1024), (4096, 64 * 3), (1000, 191)])
def test_pbtrf_wide_bulk_staging_same_bits(bm, rng, shape):
    """The update kernel stages whole slabs of U12 by bulk copies (TMA) when the band array's strides keep their 512-byte pieces
    16-byte aligned, and by 8-byte cp.async otherwise (tuning pb_nobulk = 1 forces the latter): the arithmetic is the same, so
    the factors must agree bit for bit, for both parities of the leading dimension."""
    n, kd = shape
    try:
        for uplo, extra in itertools.product("UL", (0, 1)):
            ab = _spd_band(rng, n, kd, uplo, extra)
            out = []
            for nobulk in (1, 0):
                bm.handle(0).tune("pb_nobulk", nobulk)
                dA = _dev(ab)
                _, info = bm.pbtrf_(uplo, n, kd, dA)
                assert info == 0
                out.append(_host(dA))
            assert np.array_equal(out[0], out[1]), (uplo, n, kd, extra)
    finally:
        bm.handle(0).tune("reset", 0)


@pytest.mark.parametrize("kd", [3, 7, 20, 40, 100])
def test_pbtrf_not_positive_definite(bm, oracle_c, rng, kd):
    n = 400
    for uplo in "UL":
        ab = _spd_band(rng, n, kd, uplo)
        ab[kd if uplo == "U" else 0, 217] = -1.0
        dA = _dev(ab)
        _, info = bm.pbtrf_(uplo, n, kd, dA)
        assert info == 218
        if kd <= 64:  # DPBTF2 leaves the trailing window as updated so far
            ref = ab.copy(order="F")
            assert oracle_c.pbtrf(uplo, n, kd, ref, kd + 1) == 218
            assert np.array_equal(_host(dA), ref)
        A = bm.BandedMatrix.from_banddata(np.vstack([ab, np.zeros((kd, n))]) if uplo == "U" else np.vstack([np.zeros((kd, n)), ab]), n, kd, kd)
        with pytest.raises(bm.PosDefException):
            bm.cholesky(A, uplo)
        assert bm.cholesky(A, uplo, check=False).info == 218


def test_cholesky_of_a_banded_matrix_and_solve(bm, rng):
    """cholesky(Symmetric(A)) \\ b on the 1-D Laplacian-like SPD band, both triangles, vector and matrix right-hand sides."""
    n, k = 2000, 5
    D = np.zeros((2 * k + 1, n), order="F")
    D[k] = 4.0 * k
    for d in range(1, k + 1):
        D[k - d, d:] = -1.0 / d
        D[k + d, : n - d] = -1.0 / d
    A = bm.BandedMatrix.from_banddata(D, n, k, k)
    dense = A.to_dense()
    b = rng.standard_normal((n, 4))
    xref = np.linalg.solve(dense, b)
    for uplo in "UL":
        F = bm.cholesky(A, uplo)
        assert F.issuccess()
        X = bm.solve(F, bm.to_colmajor(b)).cpu().numpy()
        assert np.max(np.abs(X - xref)) <= 1e-12 * np.max(np.abs(xref))
        x = bm.solve(F, torch.as_tensor(b[:, 1].copy()).cuda()).cpu().numpy()
        assert np.max(np.abs(x - xref[:, 1])) <= 1e-12 * np.max(np.abs(xref))
    assert np.array_equal(A.to_dense(), dense)  # cholesky copies (cholcopy)


def test_pbtrf_nan_propagates(bm, oracle_c, rng):
    n, kd = 300, 6
    ab = _spd_band(rng, n, kd, "U")
    ab[kd - 2, 150] = np.nan
    ref = ab.copy(order="F")
    iref = oracle_c.pbtrf("U", n, kd, ref, kd + 1)
    dA = _dev(ab)
    _, info = bm.pbtrf_("U", n, kd, dA)
    assert info == iref
    assert np.array_equal(_host(dA), ref, equal_nan=True)


def test_pbtrf_argument_errors(bm):
    t = torch.zeros((4, 2), dtype=torch.float64, device="cuda")
    with pytest.raises(ValueError):
        bm.pbtrf_("U", 4, 3, t)  # not enough bands
    with pytest.raises(ValueError):
        bm.pbtrf_("X", 4, 1, t)
    with pytest.raises(ValueError):
        bm.pbtrf_("U", 5, 1, t)  # not square
    with pytest.raises(bm.DimensionMismatch):
        bm.pbtrs_("U", 4, 1, t, torch.zeros(5, dtype=torch.float64, device="cuda"))
