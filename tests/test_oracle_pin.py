"""CPU: pins the oracle.  (1) the C restatement (oracle/bmoracle.c) against the OpenBLAS 0.3.30 ILP64 library
numpy ships -- the Fortran entry points the reference ccalls -- bit-for-bit where the operation order is defined
(gbmv 'N', gbtf2, gbtrs 'N') and to tolerance elsewhere; (2) both against the committed golden fixtures; (3) both
against the reference's deterministic integer KAT (test/test_linalg.jl:51-97) and dense arithmetic."""
import itertools

import numpy as np
import pytest

import oracle
from oracle import Band, band_from_dense, banded_muladd_vec, brand, gbmm_kernel, gbmv, ldiv, lu

from _util import golden_cases, kat_matrix, scalar


def test_openblas_identity(oracle_ob):
    assert "OpenBLAS 0.3.30" in oracle_ob.config and "USE64BITINT" in oracle_ob.config


@pytest.mark.parametrize("shape", [(100, 100, 4, 3), (40, 55, 2, 5), (55, 40, 5, 2), (300, 300, 32, 32), (10, 10, 0, 0),
                                   (1, 10, 0, 9), (50, 50, 1, 0), (10, 10, 12, 15)])
def test_c_gbmv_matches_openblas(oracle_c, oracle_ob, rng, shape):
    m, n, kl, ku = shape
    for (al, be) in [(1.0, 0.0), (2.0, 3.0), (0.123, 0.456), (0.0, 0.0), (0.0, 2.0), (1.0, 1.0)]:
        A = brand(rng, m, n, kl, ku, corners=np.nan)  # NaN corners: must never be read
        for tr in "NT":
            x = rng.standard_normal(n if tr == "N" else m)
            y0 = rng.standard_normal(m if tr == "N" else n)
            if be == 0.0:
                y0[:] = np.nan  # beta == 0 overwrites (test/test_linalg.jl:225-270)
            y1, y2 = y0.copy(), y0.copy()
            gbmv(oracle_c, tr, m, kl, ku, al, A.data, x, be, y1)
            gbmv(oracle_ob, tr, m, kl, ku, al, A.data, x, be, y2)
            if tr == "N":
                assert np.array_equal(y1, y2)
            else:
                assert np.allclose(y1, y2, rtol=1e-13, atol=1e-13)
            D = A.dense()
            ref = al * ((D @ x) if tr == "N" else (D.T @ x)) + (0 if be == 0 else be * y0)
            assert np.allclose(y1, ref, rtol=1e-12, atol=1e-12)


def test_c_gbmv_strided(oracle_c, oracle_ob, rng):
    m, n, kl, ku = 30, 40, 3, 2
    A = brand(rng, m, n, kl, ku)
    xs, ys = rng.standard_normal(3 * n), rng.standard_normal(2 * m)
    y1, y2 = ys.copy(), ys.copy()
    gbmv(oracle_c, "N", m, kl, ku, 1.5, A.data, xs[::3], 0.5, y1[::2])
    gbmv(oracle_ob, "N", m, kl, ku, 1.5, A.data, xs[::3], 0.5, y2[::2])
    assert np.array_equal(y1, y2)


def test_c_gbmm_regimes_match_openblas_all_shapes(oracle_c, oracle_ob, rng):
    """The reference's own gbmm! sweep (test/test_linalg.jl:212-222): 3^3 * 4^4 = 6912 shape combinations."""
    L = oracle_c.L
    cnt = 0
    for n, nu, m in itertools.product([1, 5, 50], repeat=3):
        for Al, Au, Bl, Bu in itertools.product([0, 1, 2, 30], repeat=4):
            A = brand(rng, n, nu, Al, Au, corners=np.nan)
            B = brand(rng, nu, m, Bl, Bu, corners=np.nan)
            Cl, Cu = min(n - 1, Al + Bl), min(m - 1, Au + Bu)
            C0 = brand(rng, n, m, Cl, Cu)
            c1, c2, c3 = (C0.data.copy(order="F") for _ in range(3))
            gbmm_kernel(oracle_c, 0.123, A.data, B.data, 0.456, c1, n, nu, m, Al, Au, Bl, Bu, Cl, Cu)
            gbmm_kernel(oracle_ob, 0.123, A.data, B.data, 0.456, c2, n, nu, m, Al, Au, Bl, Bu, Cl, Cu)
            L.oracle_gbmm(n, nu, m, Al, Au, Bl, Bu, Cl, Cu, 0.123, A.data.ctypes.data, A.data.shape[0],
                          B.data.ctypes.data, B.data.shape[0], 0.456, c3.ctypes.data, c3.shape[0])
            assert np.array_equal(c1, c2) and np.array_equal(c1, c3)
            if cnt % 16 == 0:
                D = 0.123 * A.dense() @ B.dense() + 0.456 * C0.dense()
                assert np.allclose(Band(c1, n, Cl, Cu).dense(), D, rtol=1e-12, atol=1e-12)
            cnt += 1
    assert cnt == 6912


@pytest.mark.parametrize("shape", [(1000, 4, 3, 1), (500, 16, 16, 3), (300, 5, 7, 5), (64, 3, 2, 2), (400, 64, 64, 2),
                                   (1, 0, 0, 1), (5, 4, 4, 1), (200, 0, 3, 2), (200, 3, 0, 2), (128, 1, 1, 4)])
def test_c_lu_solve_bit_identical_to_openblas(oracle_c, oracle_ob, rng, shape):
    n, kl, ku, nrhs = shape
    A = brand(rng, n, n, kl, ku)
    ab1, p1, i1 = lu(oracle_c, A)
    ab2, p2, i2 = lu(oracle_ob, A)
    assert i1 == i2 == 0
    assert np.array_equal(p1, p2)          # pivots bit-identical
    assert np.array_equal(ab1, ab2)        # unblocked regime: factors bit-identical
    B = np.asfortranarray(rng.standard_normal((n, nrhs)))
    for tr in "NT":
        b1, b2 = B.copy(order="F"), B.copy(order="F")
        ldiv(oracle_c, tr, ab2, p2, kl, ku, b1)
        ldiv(oracle_ob, tr, ab2, p2, kl, ku, b2)
        if tr == "N":
            assert np.array_equal(b1, b2)
        else:
            assert np.max(np.abs(b1 - b2)) <= 1e-11 * np.max(np.abs(b2))


@pytest.mark.parametrize("shape", [(300, 70, 70), (260, 100, 65), (400, 128, 128)])
def test_c_lu_blocked_regime_pivots_equal(oracle_c, oracle_ob, rng, shape):
    """ku > 64 and kl >= 32: OpenBLAS runs LAPACK's blocked DGBTRF (DTRSM+DGEMM rounding): pivots equal, factors ~."""
    n, kl, ku = shape
    A = brand(rng, n, n, kl, ku)
    ab1, p1, _ = lu(oracle_c, A)
    ab2, p2, _ = lu(oracle_ob, A)
    assert np.array_equal(p1, p2)
    assert np.max(np.abs(ab1 - ab2)) < 1e-10


def test_c_lu_singular_info(oracle_c, oracle_ob):
    A = Band(np.asfortranarray(np.zeros((3, 6))), 6, 1, 1)
    A.data[1, :] = [1, 2, 0, 4, 5, 6]  # zero diagonal at column 3, no off-diagonals -> info = 3
    _, _, i1 = lu(oracle_c, A)
    _, _, i2 = lu(oracle_ob, A)
    assert i1 == i2 == 3


def test_kat_integer_valued(oracle_c, oracle_ob):
    """test/test_linalg.jl:51-97: exact in Float64 because every operand is a small integer."""
    D, v, X = kat_matrix()
    A = band_from_dense(D, 2, 2)
    for be_ in (oracle_c, oracle_ob):
        for (al, bt) in [(1.0, 0.0), (1.0, 1.0), (0.0, 1.0), (2.0, 3.0)]:
            y = v.copy()
            banded_muladd_vec(be_, al, A, v, bt, y)
            assert np.array_equal(y, al * (D @ v) + bt * v)
        ab, ipiv, info = lu(be_, A)
        assert info == 0
        sol = np.asfortranarray(X.copy())
        ldiv(be_, "N", ab, ipiv, 2, 2, sol)
        assert np.allclose(D @ sol, X, rtol=1e-10, atol=1e-10)


def test_negative_bandwidth_driver(oracle_c, rng):
    """_banded_muladd! re-viewing (src/generic/matmul.jl:41-59); shapes of test/test_banded.jl:97-140."""
    for (m, n, l, u) in [(10, 12, 2, 3), (10, 12, -2, 2), (10, 12, 2, -2), (10, 12, 2, -3), (12, 10, -1, 1), (8, 8, 1, -1),
                         (8, 8, -2, 1)]:
        D = np.triu(np.tril(rng.standard_normal((m, n)), u), -l) if -l <= u else np.zeros((m, n))
        A = band_from_dense(D, l, u)
        x, y0 = rng.standard_normal(n), rng.standard_normal(m)
        y = y0.copy()
        banded_muladd_vec(oracle_c, 2.0, A, x, 3.0, y)
        assert np.allclose(y, 2.0 * D @ x + 3.0 * y0, rtol=1e-13, atol=1e-13), (m, n, l, u)


def test_golden_gbmv(oracle_c):
    for cid, c in golden_cases("gbmv"):
        m, l, u = (int(scalar(c[k])) for k in ("m", "l", "u"))
        tr = str(scalar(c["trans"]))
        y = c["y0"].copy()
        gbmv(oracle_c, tr, m, l, u, float(c["alpha"]), np.asfortranarray(c["data"]), c["x"], float(c["beta"]), y)
        if tr == "N":
            assert np.array_equal(y, c["y"]), cid
        else:
            assert np.allclose(y, c["y"], rtol=1e-13, atol=1e-13), cid


def test_golden_gbmm(oracle_c):
    for cid, c in golden_cases("gbmm"):
        n, nu, m, Al, Au, Bl, Bu, Cl, Cu = (int(v) for v in c["dims"])
        out = np.asfortranarray(c["C0"].copy())
        gbmm_kernel(oracle_c, 0.123, np.asfortranarray(c["A"]), np.asfortranarray(c["B"]), 0.456, out, n, nu, m, Al, Au, Bl,
                    Bu, Cl, Cu)
        assert np.array_equal(out, c["C"]), cid


def test_golden_lu(oracle_c):
    for cid, c in golden_cases("lu"):
        n, l, u, nrhs, info = (int(v) for v in c["dims"])
        A = Band(np.asfortranarray(c["data"]), n, l, u)
        ab, ipiv, inf = lu(oracle_c, A)
        assert inf == info and np.array_equal(ipiv, c["ipiv"]), cid
        if not (u > 64 and l >= 32):
            assert np.array_equal(ab, c["ab"]), cid
        X = np.asfortranarray(c["B"].copy())
        ldiv(oracle_c, "N", np.asfortranarray(c["ab"]), c["ipiv"], l, u, X)
        assert np.array_equal(X, c["X"]), cid


def _tri_band(rng, n, k, uplo, lda_extra=0):
    """Triangular-band storage with a safe diagonal and small off-diagonals (unit solves stay bounded)."""
    a = np.asfortranarray(rng.standard_normal((k + 1 + lda_extra, n))) / (2 * k + 2)
    a[k if uplo == "U" else 0, :] = (1.0 + rng.random(n)) * rng.choice([-1.0, 1.0], n)
    return a


@pytest.mark.parametrize("shape", [(1, 0), (5, 2), (40, 3), (100, 16), (257, 40), (3000, 7), (2000, 300), (50, 80), (5000, 33)])
def test_c_tbsv_tbmv_match_openblas_bit_for_bit(oracle_c, oracle_ob, rng, shape):
    """tbsv! / tbmv! (src/blas.jl:71-141), 'N': the C restatement equals OpenBLAS dtbsv_ / dtbmv_ bit for bit for
    every uplo / diag, and agrees with dense arithmetic."""
    n, k = shape
    for uplo, diag, extra in itertools.product("UL", "NU", (0, 2)):
        a = _tri_band(rng, n, k, uplo, extra)
        lda = a.shape[0]
        T = np.zeros((n, n))
        for j in range(n):
            for i in range(max(0, j - k), j + 1) if uplo == "U" else range(j, min(n, j + k + 1)):
                T[i, j] = a[(k + i - j) if uplo == "U" else (i - j), j]
        if diag == "U":
            np.fill_diagonal(T, 1.0)
        for name in ("tbsv", "tbmv"):
            x0 = rng.standard_normal(n)
            x1, x2 = x0.copy(), x0.copy()
            assert getattr(oracle_c, name)(uplo, "N", diag, n, k, a, lda, x1) == 0
            getattr(oracle_ob, name)(uplo, "N", diag, n, k, a, lda, x2)
            assert np.array_equal(x1, x2), (name, uplo, diag, n, k)
            if n <= 300:
                ref = np.linalg.solve(T, x0) if name == "tbsv" else T @ x0
                assert np.allclose(x1, ref, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("shape", [(5, 2), (100, 16), (257, 40), (3000, 7), (2000, 300), (50, 80)])
def test_c_tbsv_tbmv_transposed_match_openblas(oracle_c, oracle_ob, rng, shape):
    """'T': OpenBLAS uses its SIMD dot kernel (order unspecified) -> 1e-13, plus dense arithmetic."""
    n, k = shape
    for uplo, diag in itertools.product("UL", "NU"):
        a = _tri_band(rng, n, k, uplo, 1)
        lda = a.shape[0]
        for name in ("tbsv", "tbmv"):
            x0 = rng.standard_normal(n)
            x1, x2 = x0.copy(), x0.copy()
            assert getattr(oracle_c, name)(uplo, "T", diag, n, k, a, lda, x1) == 0
            getattr(oracle_ob, name)(uplo, "T", diag, n, k, a, lda, x2)
            assert np.max(np.abs(x1 - x2)) <= 1e-13 * max(1.0, np.max(np.abs(x2))), (name, uplo, diag)


@pytest.mark.parametrize("shape", [(1, 0), (5, 2), (40, 3), (100, 16), (257, 40), (3000, 7), (2000, 300), (50, 80)])
def test_c_sbmv_matches_openblas(oracle_c, oracle_ob, rng, shape):
    """sbmv! (src/blas.jl:36-66): the C restatement of DSBMV vs OpenBLAS dsbmv_ (axpy + SIMD dot per column: 1e-13) and
    vs dense arithmetic; beta == 0 overwrites NaN, alpha == 0 does not read A."""
    n, k = shape
    for uplo, (al, be) in itertools.product("UL", [(1.0, 0.0), (0.7, -1.3), (0.0, 2.0), (0.0, 0.0)]):
        lda = k + 1 + (n % 3)
        a = np.asfortranarray(rng.standard_normal((lda, n)))
        S = np.zeros((n, n))
        for j in range(n):
            for i in (range(max(0, j - k), j + 1) if uplo == "U" else range(j, min(n, j + k + 1))):
                S[i, j] = S[j, i] = a[(k + i - j) if uplo == "U" else (i - j), j]
        if al == 0.0:
            a[:] = np.nan
        x = rng.standard_normal(n)
        y0 = rng.standard_normal(n)
        if be == 0.0:
            y0[:] = np.nan
        y1, y2 = y0.copy(), y0.copy()
        assert oracle_c.sbmv(uplo, n, k, al, a, lda, x, be, y1) == 0
        oracle_ob.sbmv(uplo, n, k, al, a, lda, x, be, y2)
        assert np.isfinite(y1).all() and np.isfinite(y2).all()
        assert np.max(np.abs(y1 - y2)) <= 1e-13 * max(1.0, np.max(np.abs(y2)))
        if al != 0.0 and n <= 300:
            ref = al * (S @ x) + (0 if be == 0 else be * y0)
            assert np.allclose(y1, ref, rtol=1e-12, atol=1e-12)


def _spd_band(rng, n, kd, uplo, lda_extra=0):
    """SPD band matrix in LAPACK symmetric band storage (`uplo` triangle), diagonally dominant."""
    ab = np.asfortranarray(rng.standard_normal((kd + 1 + lda_extra, n)))
    ab[kd if uplo == "U" else 0, :] = 2.0 * (kd + 1) + rng.random(n)
    return ab


@pytest.mark.parametrize("shape", [(1, 0), (7, 2), (40, 3), (300, 16), (500, 31), (400, 33), (700, 64), (20, 40)])
def test_c_pbtf2_pbtrs_match_openblas_bit_for_bit(oracle_c, oracle_ob, rng, shape):
    """pbtrf! / pbtrs! (src/lapack.jl:268-332): for kd <= 64 DPBTRF runs the unblocked DPBTF2 (ILAENV: NB = 1), whose
    operation order is defined -- the C restatement equals OpenBLAS dpbtrf_ bit for bit; dpbtrs_ = two dtbsv_ sweeps."""
    n, kd = shape
    for uplo, extra in itertools.product("UL", (0, 2)):
        ab = _spd_band(rng, n, kd, uplo, extra)
        ldab = ab.shape[0]
        a1, a2 = ab.copy(order="F"), ab.copy(order="F")
        assert oracle_c.pbtrf(uplo, n, kd, a1, ldab) == 0
        assert oracle_ob.pbtrf(uplo, n, kd, a2, ldab) == 0
        rows = slice(0, kd + 1)
        assert np.array_equal(a1[rows], a2[rows]), (uplo, n, kd, extra)
        b = np.asfortranarray(rng.standard_normal((n, 3)))
        b1, b2 = b.copy(order="F"), b.copy(order="F")
        assert oracle_c.pbtrs(uplo, n, kd, 3, a1, ldab, b1, n) == 0
        assert oracle_ob.pbtrs(uplo, n, kd, 3, a2, ldab, b2, n) == 0
        assert np.max(np.abs(b1 - b2)) <= 1e-13 * max(1.0, np.max(np.abs(b2))), (uplo, n, kd)


def test_c_pbtf2_not_positive_definite(oracle_c, oracle_ob, rng):
    n, kd = 50, 4
    for uplo in "UL":
        ab = _spd_band(rng, n, kd, uplo)
        ab[kd if uplo == "U" else 0, 17] = -1.0
        a1, a2 = ab.copy(order="F"), ab.copy(order="F")
        assert oracle_c.pbtrf(uplo, n, kd, a1, kd + 1) == 18
        assert oracle_ob.pbtrf(uplo, n, kd, a2, kd + 1) == 18
        assert np.array_equal(a1, a2)


@pytest.mark.parametrize("dt", [np.float32, np.complex64, np.complex128])
def test_openblas_typed_entry_points_match_dense_numpy(oracle_ob, rng, dt):
    """The S / C / Z checker used by tests/test_gpu_typed.py (OpenBLAS {s,c,z}gbmv_ / hbmv_ / gbtrf_ / gbtrs_ entered as the
    reference does, src/blas.jl:4-66) against a dense numpy restatement of the same operations."""
    tol = 1e-4 if dt in (np.float32, np.complex64) else 1e-12
    cplx = np.issubdtype(dt, np.complexfloating)

    def rand(shape):
        a = rng.standard_normal(shape)
        return np.asfortranarray((a + 1j * rng.standard_normal(shape) if cplx else a).astype(dt))

    m, n, kl, ku = 60, 50, 4, 3
    a = rand((kl + ku + 1, n))
    D = np.zeros((m, n), dtype=dt)
    for j in range(n):
        for i in range(max(0, j - ku), min(m, j + kl + 1)):
            D[i, j] = a[ku + i - j, j]
    alpha, beta = (0.5 - 1j, 0.25 + 0.5j) if cplx else (0.5, 0.25)
    for trans, op in (("N", D), ("T", D.T), ("C", D.conj().T)):
        x, y = rand(op.shape[1]), rand(op.shape[0])
        ref = alpha * (op.astype(np.complex128 if cplx else np.float64) @ x) + beta * y
        oracle_ob.t_gbmv(trans, m, n, kl, ku, alpha, a, kl + ku + 1, x, beta, y)
        assert np.max(np.abs(y - ref)) <= tol * np.max(np.abs(ref))
    # Hermitian band
    n, k = 40, 3
    for uplo in "UL":
        h = rand((k + 1, n))
        H = np.zeros((n, n), dtype=dt)
        for j in range(n):
            for d in range(k + 1):
                i = j - d if uplo == "U" else j + d
                if 0 <= i < n:
                    v = h[k - d if uplo == "U" else d, j]
                    if d == 0:
                        H[j, j] = v.real
                    else:
                        H[i, j] = v
                        H[j, i] = np.conj(v)
        x, y = rand(n), rand(n)
        ref = alpha * (H @ x) + beta * y
        oracle_ob.t_hbmv(uplo, n, k, alpha, h, k + 1, x, beta, y)
        assert np.max(np.abs(y - ref)) <= tol * np.max(np.abs(ref))
    # LU + the three solves
    n, kl, ku = 50, 3, 2
    ldab = 2 * kl + ku + 1
    ab = rand((ldab, n))
    ab[:kl] = 0
    ab[kl + ku] += 4.0
    D = np.zeros((n, n), dtype=dt)
    for j in range(n):
        for i in range(max(0, j - ku), min(n, j + kl + 1)):
            D[i, j] = ab[kl + ku + i - j, j]
    ipiv = np.zeros(n, dtype=np.int64)
    assert oracle_ob.t_gbtrf(n, n, kl, ku, ab, ldab, ipiv) == 0
    for trans, op in (("N", D), ("T", D.T), ("C", D.conj().T)):
        b = rand((n, 2))
        ref = np.linalg.solve(op.astype(np.complex128 if cplx else np.float64), b)
        assert oracle_ob.t_gbtrs(trans, n, kl, ku, 2, ab, ldab, ipiv, b, n) == 0
        assert np.max(np.abs(b - ref)) <= 10 * tol * np.max(np.abs(ref))
