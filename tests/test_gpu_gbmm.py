"""GPU parity: mul!(C, A, B, α, β) for banded x banded (gbmm!, src/banded/gbmm.jl) and banded x dense
(src/generic/matmul.jl:243-271) vs the oracle's replay of the reference call sequence -- bit-exact.
Shapes from test/test_linalg.jl:138-325, test/test_banded.jl:142-253, test/test_broadcasting.jl:457-478."""
import itertools

import numpy as np
import pytest
import torch

import oracle
from oracle import Band, band_from_dense, brand, gbmm_kernel

from _util import golden_cases, relerr

pytestmark = pytest.mark.gpu


def up(bm, Bd: Band):
    return bm.BandedMatrix.from_banddata(Bd.data, Bd.m, Bd.l, Bd.u)


def test_golden_gbmm(bm):
    for cid, c in golden_cases("gbmm"):
        n, nu, m, Al, Au, Bl, Bu, Cl, Cu = (int(v) for v in c["dims"])
        A = bm.BandedMatrix.from_banddata(c["A"], n, Al, Au)
        B = bm.BandedMatrix.from_banddata(c["B"], nu, Bl, Bu)
        Cm = bm.BandedMatrix.from_banddata(c["C0"], n, Cl, Cu)
        bm.mul_(Cm, A, B, 0.123, 0.456)
        assert np.array_equal(Cm.banddata_host(), c["C"]), cid


def test_gbmm_reference_sweep(bm, oracle_c, rng):
    """test/test_linalg.jl:212-222, thinned: every (n,ν,m) in {1,5,50}^3, bands from {0,1,2,30}."""
    combos = list(itertools.product([0, 1, 2, 30], repeat=4))
    for n, nu, m in itertools.product([1, 5, 50], repeat=3):
        for Al, Au, Bl, Bu in combos[:: 5]:
            A = brand(rng, n, nu, Al, Au, corners=np.nan)
            B = brand(rng, nu, m, Bl, Bu, corners=np.nan)
            Cl, Cu = min(n - 1, Al + Bl), min(m - 1, Au + Bu)
            C0 = brand(rng, n, m, Cl, Cu)
            ref = C0.data.copy(order="F")
            gbmm_kernel(oracle_c, 0.123, A.data, B.data, 0.456, ref, n, nu, m, Al, Au, Bl, Bu, Cl, Cu)
            Cm = up(bm, C0)
            bm.mul_(Cm, up(bm, A), up(bm, B), 0.123, 0.456)
            assert np.array_equal(Cm.banddata_host(), ref), (n, nu, m, Al, Au, Bl, Bu)


@pytest.mark.parametrize("shape", [(1000, 1000, 1000, 4, 3, 4, 3), (600, 500, 700, 32, 32, 32, 32), (300, 300, 300, 5, 0, 0, 7),
                                   (2000, 2000, 2000, 1, 1, 2, 2), (400, 450, 380, 64, 10, 3, 40)])
def test_gbmm_larger_and_star(bm, oracle_c, rng, shape):
    n, nu, m, Al, Au, Bl, Bu = shape
    A, B = brand(rng, n, nu, Al, Au, corners=np.nan), brand(rng, nu, m, Bl, Bu, corners=np.nan)
    Cl, Cu = min(n - 1, Al + Bl), min(m - 1, Au + Bu)
    ref = np.full((Cl + Cu + 1, m), np.nan, order="F")  # β = 0 must overwrite NaN (test_linalg.jl:225-270)
    gbmm_kernel(oracle_c, 1.0, A.data, B.data, 0.0, ref, n, nu, m, Al, Au, Bl, Bu, Cl, Cu)
    P = bm.matmul(up(bm, A), up(bm, B))  # A*B allocates (Cl,Cu) = min.(size-1, sums)  (matmul.jl:1-6)
    assert (P.l, P.u) == (Cl, Cu)
    got = P.banddata_host()
    # in-matrix entries bit-identical; out-of-matrix corner slots of an undef C are unspecified
    assert np.array_equal(Band(got, n, Cl, Cu).dense(), Band(ref, n, Cl, Cu).dense())
    assert np.allclose(P.to_dense(), A.dense() @ B.dense(), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("shape", [(3000, 3000, 3000, 32, 32, 32, 32), (2500, 2400, 2600, 40, 9, 12, 30), (1500, 1500, 1500, 8, 8, 8, 8),
                                   (2100, 2000, 1900, 64, 64, 20, 20)])
def test_gbmm_tensor_core_path_multi_tile(bm, oracle_c, rng, shape):
    """Many column tiles per CTA on the DMMA kernel, alpha/beta both non-trivial: bit-identical to the per-column dgbmv_ replay."""
    n, nu, m, Al, Au, Bl, Bu = shape
    A, B = brand(rng, n, nu, Al, Au, corners=np.nan), brand(rng, nu, m, Bl, Bu, corners=np.nan)
    Cl, Cu = min(n - 1, Al + Bl), min(m - 1, Au + Bu)
    C0 = brand(rng, n, m, Cl, Cu)
    ref = C0.data.copy(order="F")
    gbmm_kernel(oracle_c, -0.75, A.data, B.data, 1.25, ref, n, nu, m, Al, Au, Bl, Bu, Cl, Cu)
    Cm = up(bm, C0)
    bm.mul_(Cm, up(bm, A), up(bm, B), -0.75, 1.25)
    assert np.array_equal(Band(Cm.banddata_host(), n, Cl, Cu).dense(), Band(ref, n, Cl, Cu).dense())


@pytest.mark.parametrize("shape", [(3000, 3000, 3000, 64, 64, 64, 64), (2500, 2500, 2500, 48, 48, 48, 48), (2100, 2000, 2300, 100, 40, 50, 90),
                                   (5000, 5000, 5000, 40, 60, 70, 30)])
def test_gbmm_ring_kernel_wide_bands(bm, oracle_c, rng, shape):
    """Bands too wide for the two-CTA tile kernel: the persistent ring kernel (gbmm_bb_ring: staged A columns kept in a
    shared-memory ring across consecutive column tiles, 8/16/32-column tiles by fit).  Same per-element order => same bits."""
    n, nu, m, Al, Au, Bl, Bu = shape
    A, B = brand(rng, n, nu, Al, Au, corners=np.nan), brand(rng, nu, m, Bl, Bu, corners=np.nan)
    Cl, Cu = min(n - 1, Al + Bl), min(m - 1, Au + Bu)
    for alpha, beta in [(1.0, 0.0), (-0.75, 1.25)]:
        C0 = brand(rng, n, m, Cl, Cu)
        ref = C0.data.copy(order="F")
        gbmm_kernel(oracle_c, alpha, A.data, B.data, beta, ref, n, nu, m, Al, Au, Bl, Bu, Cl, Cu)
        Cm = up(bm, C0)
        bm.mul_(Cm, up(bm, A), up(bm, B), alpha, beta)
        assert np.array_equal(Band(Cm.banddata_host(), n, Cl, Cu).dense(), Band(ref, n, Cl, Cu).dense())


@pytest.mark.parametrize("shape", [(2000, 2000, 2000, 72, 70, 72, 70), (1500, 1400, 1600, 100, 90, 80, 120), (1200, 1200, 1200, 300, 300, 300, 300),
                                   (2500, 2500, 2500, 1024, 1024, 1024, 1024), (900, 1000, 800, 200, 10, 5, 400), (300, 300, 300, 299, 299, 299, 299),
                                   (700, 650, 720, 12, 9, 10, 14)])
def test_gbmm_kblocked_wide_band_kernel(bm, oracle_c, rng, shape):
    """Bands too wide for staged whole columns: the K-blocked tensor-core kernel (gbmm_wide.cu; the last shape forces it on a
    narrow band through the tuning block).  Same per-element order => bit-identical to the per-column dgbmv_ replay."""
    n, nu, m, Al, Au, Bl, Bu = shape
    A, B = brand(rng, n, nu, Al, Au, corners=np.nan), brand(rng, nu, m, Bl, Bu, corners=np.nan)
    Cl, Cu = min(n - 1, Al + Bl), min(m - 1, Au + Bu)
    hd = bm.handle(0)
    hd.tune("gbmm_wide", 1)
    try:
        for alpha, beta in [(1.0, 0.0), (-0.75, 1.25)]:
            C0 = brand(rng, n, m, Cl, Cu)
            ref = C0.data.copy(order="F")
            gbmm_kernel(oracle_c, alpha, A.data, B.data, beta, ref, n, nu, m, Al, Au, Bl, Bu, Cl, Cu)
            Cm = up(bm, C0)
            l0 = hd.launches
            bm.mul_(Cm, up(bm, A), up(bm, B), alpha, beta)
            assert hd.launches > l0 and hd.last_gbmm_path() == 3
            assert np.array_equal(Band(Cm.banddata_host(), n, Cl, Cu).dense(), Band(ref, n, Cl, Cu).dense())
    finally:
        hd.tune("reset", 0)


def test_gbmm_wider_C_and_banderror(bm, rng):
    """C with extra bands gets zeros/β-scaling there (test_broadcasting.jl:457-478); too few bands ⇒ BandError
    unless the missing bands are structurally zero (test_linalg.jl:272-295)."""
    n = 20
    DA = np.triu(np.tril(rng.standard_normal((n, n)), 1), -1)
    DB = np.triu(np.tril(rng.standard_normal((n, n)), 2), -2)
    A, B = bm.BandedMatrix.from_dense(DA, (1, 1)), bm.BandedMatrix.from_dense(DB, (2, 2))
    for (cl, cu) in [(3, 3), (4, 4), (5, 3)]:
        C0 = np.triu(np.tril(rng.standard_normal((n, n)), cu), -cl)
        Cm = bm.BandedMatrix.from_dense(C0, (cl, cu))
        bm.mul_(Cm, A, B, 2.0, 3.0)
        assert np.allclose(Cm.to_dense(), 2.0 * DA @ DB + 3.0 * C0, rtol=1e-13, atol=1e-13)
        Cn = bm.BandedMatrix.from_banddata(np.full((cl + cu + 1, n), np.nan), n, cl, cu)
        bm.mul_(Cn, A, B)
        assert np.allclose(Cn.to_dense(), DA @ DB, rtol=1e-13, atol=1e-13)
    with pytest.raises(bm.BandError):
        bm.mul_(bm.BandedMatrix.zeros((n, n), (2, 2)), A, B)
    # B's outer bands are zero: a (2,2) destination is then legal
    DB2 = np.triu(np.tril(DB, 1), -1)
    B2 = bm.BandedMatrix.from_dense(DB2, (2, 2))
    C2 = bm.BandedMatrix.zeros((n, n), (2, 2))
    bm.mul_(C2, A, B2)
    assert np.allclose(C2.to_dense(), DA @ DB2, rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("bands", [(-1, 2, 1, 1), (2, -1, 1, 1), (1, 1, -1, 2), (1, 1, 2, -1), (-3, 2, 1, 1), (0, 0, 1, -2)])
def test_gbmm_negative_bands(bm, rng, bands):
    """gbmm.jl:231-249 pruning branches (test_banded.jl:165-253 'negative bands')."""
    Al, Au, Bl, Bu = bands
    n = 12
    DA = np.triu(np.tril(rng.standard_normal((n, n)), Au), -Al) if -Al <= Au else np.zeros((n, n))
    DB = np.triu(np.tril(rng.standard_normal((n, n)), Bu), -Bl) if -Bl <= Bu else np.zeros((n, n))
    A, B = bm.BandedMatrix.from_dense(DA, (Al, Au)), bm.BandedMatrix.from_dense(DB, (Bl, Bu))
    P = bm.matmul(A, B)
    assert np.allclose(P.to_dense(), DA @ DB, rtol=1e-13, atol=1e-13)


def test_transposed_operands(bm, rng):
    n = 30
    DA = np.triu(np.tril(rng.standard_normal((n, n)), 3), -1)
    DB = np.triu(np.tril(rng.standard_normal((n, n)), 0), -2)
    A, B = bm.BandedMatrix.from_dense(DA, (1, 3)), bm.BandedMatrix.from_dense(DB, (2, 0))
    assert np.allclose(bm.matmul(A.T, B).to_dense(), DA.T @ DB, rtol=1e-13, atol=1e-13)
    assert np.allclose(bm.matmul(A, B.T).to_dense(), DA @ DB.T, rtol=1e-13, atol=1e-13)
    assert np.allclose(bm.matmul(A.T, B.T).to_dense(), DA.T @ DB.T, rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("shape", [(1000, 1000, 200, 300, 7), (1200, 1000, 30, 20, 33), (500, 500, 4, 3, 256), (64, 80, 0, 0, 3)])
def test_banded_times_dense(bm, oracle_c, rng, shape):
    """test_banded.jl:142-163: every column equals the reference's per-column gbmv, bit for bit."""
    m, n, l, u, nrhs = shape
    A = brand(rng, m, n, l, u, corners=np.nan)
    X = np.asfortranarray(rng.standard_normal((n, nrhs)))
    C0 = np.asfortranarray(rng.standard_normal((m, nrhs)))
    ref = C0.copy(order="F")
    for c in range(nrhs):
        oracle.gbmv(oracle_c, "N", m, l, u, 0.7, A.data, X[:, c], 1.3, ref[:, c])
    Cd = bm.to_colmajor(C0)
    bm.mul_(Cd, up(bm, A), bm.to_colmajor(X), 0.7, 1.3)
    assert np.array_equal(Cd.cpu().numpy(), ref)
    Y = np.asfortranarray(rng.standard_normal((m, nrhs)))
    got = bm.matmul(up(bm, A).T, bm.to_colmajor(Y)).cpu().numpy()
    assert np.allclose(got, A.dense().T @ Y, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("shape", [(37, 300, 4, 3), (5, 64, 0, 2), (128, 1000, 33, 17), (1, 9, 1, 1), (200, 50, 60, 70)])
def test_dense_times_banded(bm, oracle_c, rng, shape):
    """materialize!(MatMulMatAdd{Strided,BandedColumns,Strided}) (src/generic/matmul.jl:258-271): the reference loops one strided
    gbmv('T') per ROW of C; bmb200_dgbmm_db does all rows in one launch.  Dense * banded is a dot product per entry (1e-13);
    dense * transpose(banded) is the row-wise dgbmv_('N') and must be bit-identical to it."""
    M, n, l, u = shape
    B = brand(rng, n, n, l, u)
    Ad = np.asfortranarray(rng.standard_normal((M, n)))
    C0 = np.asfortranarray(rng.standard_normal((M, n)))
    Bd = up(bm, B)
    for alpha, beta in [(1.0, 0.0), (-0.75, 1.5)]:
        # dense * banded
        ref = np.empty((M, n))
        for i in range(M):
            y = C0[i].copy()
            oracle.gbmv(oracle_c, "T", n, l, u, alpha, B.data, Ad[i].copy(), beta, y)
            ref[i] = y
        Cd = bm.to_colmajor(C0)
        bm.mul_(Cd, bm.to_colmajor(Ad), Bd, alpha, beta)
        assert relerr(Cd.cpu().numpy(), ref) <= 1e-13
        # dense * transpose(banded): rows are dgbmv_('N')
        refT = np.empty((M, n))
        for i in range(M):
            y = C0[i].copy()
            oracle.gbmv(oracle_c, "N", n, l, u, alpha, B.data, Ad[i].copy(), beta, y)
            refT[i] = y
        CdT = bm.to_colmajor(C0)
        bm.mul_(CdT, bm.to_colmajor(Ad), Bd.T, alpha, beta)
        assert np.array_equal(CdT.cpu().numpy(), refT)
    Cn = bm.to_colmajor(np.full((M, n), np.nan))
    bm.mul_(Cn, bm.to_colmajor(Ad), Bd, 1.0, 0.0)  # beta == 0 overwrites NaN
    assert np.isfinite(Cn.cpu().numpy()).all()
