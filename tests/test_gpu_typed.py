"""GPU parity of the Float32 / ComplexF32 / ComplexF64 instantiations (src/blas.jl:4-7: gbmv! / sbmv! / hbmv!; LAPACK.gbtrf! /
gbtrs! for the four element types, BandedLU.jl:98, linalg.jl:28,46,62 incl. the conjugate-transpose solve) through the C ABI
against the OpenBLAS entry points the reference ccalls ({s,c,z}gbmv_, {c,z}hbmv_, ssbmv_, {s,c,z}gbtrf_ / gbtrs_).  OpenBLAS'
operation order is unspecified for these kernels: agreement to rounding (1e-5 single, 1e-13 double, relative to the result's
scale), pivots equal (random columns have a unique maximum)."""
import itertools

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DTYPES = [np.float32, np.complex64, np.complex128]
TOL = {np.float32: 2e-5, np.complex64: 2e-5, np.complex128: 1e-13, np.float64: 1e-13}


def _rand(rng, shape, dt):
    a = rng.standard_normal(shape)
    if np.issubdtype(dt, np.complexfloating):
        a = a + 1j * rng.standard_normal(shape)
    return np.asfortranarray(a.astype(dt))


def _dev(a):
    """(rows x n) Fortran band array -> the package's (n, rows) tensor with the same memory layout."""
    return torch.as_tensor(np.ascontiguousarray(a.T)).cuda()


def _host(t):
    return np.asfortranarray(t.cpu().numpy().T)


def _close(got, ref, dt):
    scale = max(1.0, float(np.max(np.abs(ref)))) if ref.size else 1.0
    return float(np.max(np.abs(got - ref))) <= TOL[dt] * scale if ref.size else True


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", [(1, 1, 0, 0), (10, 10, 2, 1), (300, 300, 4, 3), (500, 420, 7, 0), (400, 520, 0, 5), (2000, 2000, 40, 33),
                                   (50, 50, 60, 70), (100000, 100000, 4, 3)])
def test_typed_gbmv(bm, oracle_ob, rng, dt, shape):
    m, n, kl, ku = shape
    kle, kue = min(kl, max(m - 1, 0)), min(ku, max(n - 1, 0))
    a = _rand(rng, (kl + ku + 1, n), dt)
    dA = _dev(a)
    alpha, beta = (0.75 - 0.5j, -1.25 + 0.25j) if np.issubdtype(dt, np.complexfloating) else (0.75, -1.25)
    for trans in "NTC":
        lx, ly = (n, m) if trans == "N" else (m, n)
        x, y0 = _rand(rng, lx, dt), _rand(rng, ly, dt)
        for al, be in ((alpha, beta), (1.0, 0.0)):
            ref = y0.copy()
            oracle_ob.t_gbmv(trans, m, n, kl, ku, al, a, kl + ku + 1, x, be, ref)
            y = torch.as_tensor(y0.copy()).cuda()
            if be == 0.0:
                y.fill_(float("nan"))  # beta == 0 overwrites
            bm.gbmv_(trans, m, kl, ku, al, dA, torch.as_tensor(x).cuda(), be, y)
            assert _close(y.cpu().numpy(), ref, dt), (dt, trans, shape)
    del kle, kue


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", [(1, 0), (30, 3), (1000, 5), (700, 40), (40, 60), (50000, 3)])
def test_typed_hbmv(bm, oracle_ob, rng, dt, shape):
    n, k = shape
    alpha, beta = (0.5 + 0.25j, -0.75 + 1j) if np.issubdtype(dt, np.complexfloating) else (0.5, -0.75)
    for uplo in "UL":
        a = _rand(rng, (k + 1, n), dt)
        x, y0 = _rand(rng, n, dt), _rand(rng, n, dt)
        ref = y0.copy()
        oracle_ob.t_hbmv(uplo, n, k, alpha, a, k + 1, x, beta, ref)
        y = torch.as_tensor(y0.copy()).cuda()
        bm.hbmv_(uplo, k, alpha, _dev(a), torch.as_tensor(x).cuda(), beta, y)
        assert _close(y.cpu().numpy(), ref, dt), (dt, uplo, shape)


@pytest.mark.parametrize("dt", DTYPES + [np.float64])
@pytest.mark.parametrize("shape", [(1, 1, 0, 0), (12, 12, 2, 1), (300, 300, 4, 3), (250, 250, 1, 6), (400, 400, 16, 16), (120, 120, 40, 33),
                                   (60, 60, 70, 80), (3000, 3000, 2, 2)])
def test_typed_lu_and_solve(bm, oracle_ob, rng, dt, shape):
    """Both LU kernels (shared-memory window for the narrowest bands, global-memory otherwise; tuning key typed_nowin forces the latter)."""
    for nowin in (0, 1):
        bm.handle(0).tune("typed_nowin", nowin)
        try:
            _typed_lu_case(bm, oracle_ob, rng, dt, shape)
        finally:
            bm.handle(0).tune("reset", 0)


def _typed_lu_case(bm, oracle_ob, rng, dt, shape):
    m, n, kl, ku = shape
    ldab = 2 * kl + ku + 1
    ab = _rand(rng, (ldab, n), dt)
    ab[:kl] = 0
    ab[kl + ku] += 3.0  # keep the factorisation well conditioned (pivoting still happens)
    ref = ab.copy(order="F")
    ipiv = np.zeros(min(m, n), dtype=np.int64)
    iref = oracle_ob.t_gbtrf(m, n, kl, ku, ref, ldab, ipiv)
    dAB = _dev(ab)
    if dt is np.float64:  # the generic kernels instantiated for double, cross-checked against the tuned path's reference
        hd = bm.handle(0)
        d_ipiv = torch.empty(min(m, n), dtype=torch.int64, device="cuda")
        import ctypes as C

        info = C.c_int(0)
        hd.check(hd.lib.bmb200_internal_dgbtrf_generic(hd.h, m, n, kl, ku, C.c_void_p(dAB.data_ptr()), ldab, C.c_void_p(d_ipiv.data_ptr()), C.byref(info)),
                 "dgbtrf_generic")
        info = info.value
    else:
        _, d_ipiv, info = bm.gbtrf_(m, kl, ku, dAB)
    assert info == iref == 0
    assert np.array_equal(d_ipiv.cpu().numpy(), ipiv), (dt, shape)
    got = _host(dAB)
    # in-matrix entries of the factor band (kl+ku superdiagonals of U, kl multipliers)
    mask = np.zeros((ldab, n), dtype=bool)
    for j in range(n):
        lo, hi = max(0, j - kl - ku), min(m - 1, j + kl)
        mask[kl + ku + lo - j: kl + ku + hi - j + 1, j] = True
    scale = float(np.max(np.abs(ref[mask])))
    assert float(np.max(np.abs(got[mask] - ref[mask]))) <= 50 * TOL[dt] * scale, (dt, shape)
    for trans in "NTC":
        b = _rand(rng, (n, 3), dt)
        bref = b.copy(order="F")
        assert oracle_ob.t_gbtrs(trans, n, kl, ku, 3, ref, ldab, ipiv, bref, n) == 0
        dB = torch.as_tensor(np.ascontiguousarray(b.T)).cuda().T  # column-major (n, 3)
        if dt is np.float64:
            hd.check(hd.lib.bmb200_internal_dgbtrs_generic(hd.h, trans.encode(), n, kl, ku, 3, C.c_void_p(dAB.data_ptr()), ldab, C.c_void_p(d_ipiv.data_ptr()),
                                                           C.c_void_p(dB.data_ptr()), n), "dgbtrs_generic")
        else:
            bm.gbtrs_(trans, kl, ku, n, dAB, d_ipiv, dB)
        assert float(np.max(np.abs(dB.cpu().numpy() - bref))) <= 200 * TOL[dt] * max(1.0, float(np.max(np.abs(bref)))), (dt, trans, shape)
        x = torch.as_tensor(b[:, 0].copy()).cuda()
        if dt is not np.float64:
            bm.gbtrs_(trans, kl, ku, n, dAB, d_ipiv, x)
            assert np.array_equal(x.cpu().numpy(), dB.cpu().numpy()[:, 0])


def test_typed_singular_info_and_type_errors(bm):
    AB = torch.zeros((6, 4), dtype=torch.complex64, device="cuda")
    AB[:, 2] = 1.0
    AB[3, 2] = 0.0  # exactly zero pivot in column 4 (1-based): U(4,4) = 0
    _, _, info = bm.gbtrf_(6, 1, 1, AB)
    assert info == 4
    with pytest.raises(TypeError):
        bm.gbmv_("N", 4, 1, 1, 1.0, torch.zeros((4, 3), dtype=torch.float16, device="cuda"), torch.zeros(4, device="cuda"), 0.0, torch.zeros(4, device="cuda"))
    with pytest.raises(TypeError):
        bm.gbmv_("N", 4, 1, 1, 1.0, torch.zeros((4, 3), dtype=torch.float32, device="cuda"), torch.zeros(4, dtype=torch.float64, device="cuda"), 0.0,
                 torch.zeros(4, dtype=torch.float32, device="cuda"))


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", [(1, 0), (6, 2), (300, 3), (500, 17), (40, 60), (3000, 5)])
def test_typed_tbsv_tbmv(bm, oracle_ob, rng, dt, shape):
    """tbsv! / tbmv! (src/blas.jl:71-141) for S / C / Z, all of uplo x trans x diag, against OpenBLAS' own entry points."""
    n, k = shape
    for uplo, trans, diag in itertools.product("UL", "NTC", "NU"):
        a = _rand(rng, (k + 1, n), dt) / (2 * k + 2)
        a[k if uplo == "U" else 0, :] = (1.5 + rng.random(n)).astype(dt)
        dA = _dev(a)
        for name, fn in (("tbsv", bm.tbsv_), ("tbmv", bm.tbmv_)):
            x0 = _rand(rng, n, dt)
            ref = x0.copy()
            oracle_ob.t_tb(name, uplo, trans, diag, n, k, a, k + 1, ref)
            x = torch.as_tensor(x0.copy()).cuda()
            fn(uplo, trans, diag, n, k, dA, x)
            assert _close(x.cpu().numpy(), ref, dt) or float(np.max(np.abs(x.cpu().numpy() - ref))) <= 20 * TOL[dt] * max(1.0, float(np.max(np.abs(ref)))), (
                dt, name, uplo, trans, diag, shape)


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", [(1, 0), (8, 2), (300, 4), (400, 33), (30, 40), (2000, 3)])
def test_typed_pbtrf_pbtrs(bm, oracle_ob, rng, dt, shape):
    """pbtrf! / pbtrs! (src/lapack.jl:268-332) for S / C / Z: Hermitian positive definite bands, both triangles."""
    n, kd = shape
    for uplo in "UL":
        ab = _rand(rng, (kd + 1, n), dt)
        ab[kd if uplo == "U" else 0, :] = (2.0 * (kd + 1) + rng.random(n)).astype(dt)  # real, dominant diagonal
        ref = ab.copy(order="F")
        assert oracle_ob.t_pbtrf(uplo, n, kd, ref, kd + 1) == 0
        dA = _dev(ab)
        _, info = bm.pbtrf_(uplo, n, kd, dA)
        assert info == 0
        got = _host(dA)
        mask = np.zeros((kd + 1, n), dtype=bool)
        for d in range(kd + 1):
            if uplo == "U":
                mask[kd - d, d:] = True
            else:
                mask[d, : n - d] = True
        assert float(np.max(np.abs(got[mask] - ref[mask]))) <= 20 * TOL[dt] * float(np.max(np.abs(ref[mask]))), (dt, uplo, shape)
        b = _rand(rng, (n, 2), dt)
        bref = b.copy(order="F")
        assert oracle_ob.t_pbtrs(uplo, n, kd, 2, ref, kd + 1, bref, n) == 0
        dB = torch.as_tensor(np.ascontiguousarray(b.T)).cuda().T
        bm.pbtrs_(uplo, n, kd, dA, dB)
        assert float(np.max(np.abs(dB.cpu().numpy() - bref))) <= 50 * TOL[dt] * max(1.0, float(np.max(np.abs(bref)))), (dt, uplo, shape)
    # not positive definite
    ab = _rand(rng, (kd + 1, n), dt)
    ab[kd, :] = (2.0 * (kd + 1)).real if False else 2.0 * (kd + 1)
    if n > 3:
        ab[kd, n // 2] = -1.0
        _, info = bm.pbtrf_("U", n, kd, _dev(ab))
        assert info == n // 2 + 1


def _band_dense(a, m, n, l, u):
    D = np.zeros((m, n), dtype=a.dtype)
    for j in range(n):
        for i in range(max(0, j - u), min(m, j + l + 1)):
            D[i, j] = a[u + i - j, j]
    return D


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", [(40, 40, 40, 2, 1, 1, 3), (120, 100, 90, 4, 3, 5, 2), (60, 70, 80, 0, 3, 2, 0), (200, 200, 200, 17, 9, 12, 20)])
def test_typed_gbmm(bm, rng, dt, shape):
    """banded x banded (_gbmm!, src/banded/gbmm.jl:296-340) and banded x dense (src/generic/matmul.jl:243-256) for S / C / Z against a
    dense numpy product in the next higher precision."""
    n, nu, m, Al, Au, Bl, Bu = shape
    hi = np.complex128 if np.issubdtype(dt, np.complexfloating) else np.float64
    alpha, beta = (0.75 - 0.5j, -0.5 + 0.25j) if np.issubdtype(dt, np.complexfloating) else (0.75, -0.5)
    a, b = _rand(rng, (Al + Au + 1, nu), dt), _rand(rng, (Bl + Bu + 1, m), dt)
    Cl, Cu = min(n - 1, Al + Bl), min(m - 1, Au + Bu)
    c0 = _rand(rng, (Cl + Cu + 1, m), dt)
    DA, DB, DC = _band_dense(a, n, nu, Al, Au).astype(hi), _band_dense(b, nu, m, Bl, Bu).astype(hi), _band_dense(c0, n, m, Cl, Cu).astype(hi)
    ref = alpha * (DA @ DB) + beta * DC
    dC = _dev(c0)
    bm.gbmm_typed_(alpha, _dev(a), _dev(b), beta, dC, (n, nu, m), (Al, Au), (Bl, Bu), (Cl, Cu))
    got = _band_dense(_host(dC), n, m, Cl, Cu)
    inband = _band_dense(np.ones((Cl + Cu + 1, m), dtype=dt), n, m, Cl, Cu) != 0
    assert float(np.max(np.abs(got[inband] - ref[inband]))) <= 20 * TOL[dt] * float(np.max(np.abs(ref))), (dt, shape)
    # banded x dense, all three ops
    for trans, op in (("N", DA), ("T", DA.T), ("C", DA.conj().T)):
        X = _rand(rng, (op.shape[1], 5), dt)
        Y0 = _rand(rng, (op.shape[0], 5), dt)
        refy = alpha * (op @ X.astype(hi)) + beta * Y0.astype(hi)
        dX = torch.as_tensor(np.ascontiguousarray(X.T)).cuda().T
        dY = torch.as_tensor(np.ascontiguousarray(Y0.T)).cuda().T
        bm.gbmm_bd_typed_(trans, n, Al, Au, alpha, _dev(a), dX, beta, dY)
        assert float(np.max(np.abs(dY.cpu().numpy() - refy))) <= 20 * TOL[dt] * float(np.max(np.abs(refy))), (dt, trans, shape)
