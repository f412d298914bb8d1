"""GPU parity: mul!(y, A, x, α, β) through the C ABI vs the oracle (bit-exact for 'N'), the golden fixtures,
and the reference's own test shapes (test/test_broadcasting.jl:353-394, test/test_banded.jl:97-140,
test/test_linalg.jl:51-97, 225-270, 333-338, 475-480)."""
import numpy as np
import pytest
import torch

import oracle
from oracle import band_from_dense, banded_muladd_vec, brand, gbmv

from _util import golden_cases, kat_matrix, scalar

pytestmark = pytest.mark.gpu
TOL = 1e-13  # north_star: products within 1e-13 elementwise (relative to max(1,|ref|)); 'N' is checked for bits


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def run_mul(bm, trans, A_host, x, alpha, beta, y0):
    A = bm.BandedMatrix.from_banddata(A_host.data, A_host.m, A_host.l, A_host.u)
    y = dev(y0)
    bm.mul_(y, A.T if trans == "T" else A, dev(x), alpha, beta)
    torch.cuda.synchronize()
    return y.cpu().numpy()


def test_golden_gbmv(bm):
    for cid, c in golden_cases("gbmv"):
        m, n, l, u = (int(scalar(c[k])) for k in ("m", "n", "l", "u"))
        tr = str(scalar(c["trans"]))
        A = oracle.Band(np.asfortranarray(c["data"]), m, l, u)
        y = run_mul(bm, tr, A, c["x"], float(c["alpha"]), float(c["beta"]), c["y0"])
        if tr == "N":
            assert np.array_equal(y, c["y"]), cid  # bit-identical to OpenBLAS dgbmv_n
        else:
            assert np.max(np.abs(y - c["y"]) / np.maximum(1, np.abs(c["y"]))) <= TOL, cid


@pytest.mark.parametrize("shape", [
    (100, 100, 1, 1), (100, 100, 0, 1), (100, 100, 1, 0), (100, 100, 0, 0),      # test_broadcasting.jl:353-394
    (10, 12, 2, 3), (12, 10, 3, 2), (1000, 1000, 4, 3), (1000, 1000, 3, 4),       # README shape
    (777, 777, 7, 0), (333, 400, 0, 7), (4097, 4097, 2, 2), (513, 513, 8, 7),     # every narrow width / odd lda
    (300, 300, 32, 32), (200, 260, 20, 11), (1000, 1000, 200, 300), (1200, 1000, 100, 60),  # wide (test_banded.jl:142-163)
    (1, 10, 0, 9), (10, 1, 9, 0), (5, 5, 10, 12),                                # bandwidth >> size (test_miscs.jl:119-123)
])
def test_gbmv_matches_oracle(bm, oracle_c, rng, shape):
    m, n, l, u = shape
    A = brand(rng, m, n, l, u, corners=np.nan)  # NaN in the out-of-matrix slots: must never be read
    for (al, be) in [(1.0, 0.0), (2.0, 3.0), (0.123, 0.456), (1.0, 1.0)]:
        for tr in "NT":
            x = rng.standard_normal(n if tr == "N" else m)
            y0 = rng.standard_normal(m if tr == "N" else n)
            if be == 0.0:
                y0[:] = np.nan  # β == 0 must overwrite NaN (test_linalg.jl:225-270)
            ref = y0.copy()
            gbmv(oracle_c, tr, m, l, u, al, A.data, x, be, ref)
            got = run_mul(bm, tr, A, x, al, be, y0)
            assert not np.isnan(got).any()
            if tr == "N":
                assert np.array_equal(got, ref), (shape, al, be)
            else:
                assert np.max(np.abs(got - ref) / np.maximum(1, np.abs(ref))) <= TOL, (shape, al, be)


def test_gbmv_alpha_zero_only_scales(bm, rng):
    """α = 0 ⇒ y ← β·y exactly, A and x not referenced (test_linalg.jl:475-480)."""
    A = brand(rng, 50, 50, 2, 2)
    A.data[:] = np.nan
    y0 = rng.standard_normal(50)
    got = run_mul(bm, "N", A, np.full(50, np.nan), 0.0, 2.5, y0)
    assert np.array_equal(got, 2.5 * y0)
    got = run_mul(bm, "N", A, np.full(50, np.nan), 0.0, 0.0, np.full(50, np.nan))
    assert np.array_equal(got, np.zeros(50))


@pytest.mark.parametrize("shape", [(10, 12, 2, 3), (10, 12, -2, 2), (10, 12, 2, -2), (10, 12, 2, -3), (12, 10, -1, 1),
                                   (8, 8, 1, -1), (8, 8, -2, 1), (100, 100, -1, 1), (100, 100, 1, -1), (100, 100, -2, 1)])
def test_negative_bandwidths(bm, oracle_c, rng, shape):
    """_banded_muladd! re-viewing (matmul.jl:41-59) incl. A'*x (matmul.jl:66-92): test_banded.jl:97-140."""
    m, n, l, u = shape
    D = np.triu(np.tril(rng.standard_normal((m, n)), u), -l) if -l <= u else np.zeros((m, n))
    Ah = band_from_dense(D, l, u)
    A = bm.BandedMatrix.from_banddata(Ah.data, m, l, u)
    x, y0 = rng.standard_normal(n), rng.standard_normal(m)
    y = dev(y0)
    bm.mul_(y, A, dev(x), 2.0, 3.0)
    ref = y0.copy()
    banded_muladd_vec(oracle_c, 2.0, Ah, x, 3.0, ref)
    assert np.array_equal(y.cpu().numpy(), ref)
    assert np.allclose(ref, 2.0 * D @ x + 3.0 * y0, rtol=1e-13, atol=1e-13)
    w, z0 = rng.standard_normal(m), rng.standard_normal(n)
    z = dev(z0)
    bm.mul_(z, A.T, dev(w), 2.0, 3.0)
    assert np.allclose(z.cpu().numpy(), 2.0 * D.T @ w + 3.0 * z0, rtol=1e-13, atol=1e-13)
    assert np.allclose(bm.matmul(A, dev(x)).cpu().numpy(), D @ x, rtol=1e-13, atol=1e-13)


def test_kat_and_strided_views(bm):
    """test/test_linalg.jl:51-97 (integer-valued, exact) with strided x / y views."""
    D, v, X = kat_matrix()
    A = bm.BandedMatrix.from_dense(D, (2, 2))
    for (al, be) in [(1.0, 0.0), (1.0, 1.0), (0.0, 1.0), (2.0, 3.0)]:
        y = dev(v.copy())
        bm.mul_(y, A, dev(v), al, be)
        assert np.array_equal(y.cpu().numpy(), al * (D @ v) + be * v)
    big = dev(np.arange(1.0, 31.0))
    ybig = dev(np.zeros(20))
    bm.mul_(ybig[::2], A, big[::3], 1.0, 0.0)  # non-unit incx / incy
    assert np.array_equal(ybig.cpu().numpy()[::2], D @ np.arange(1.0, 31.0)[::3])
    assert np.array_equal(ybig.cpu().numpy()[1::2], np.zeros(10))
    Xd = bm.to_colmajor(X)
    assert np.array_equal(bm.matmul(A, Xd).cpu().numpy(), D @ X)      # banded * dense
    assert np.array_equal(bm.matmul(Xd, A).cpu().numpy(), X @ D)      # dense * banded
    assert np.array_equal(bm.matmul(A.T, Xd).cpu().numpy(), D.T @ X)


def test_empty_and_mismatch(bm):
    """Empty 10x0 / 0x10 (test_banded.jl:132-139) and DimensionMismatch (test_linalg.jl:333-338)."""
    A = bm.brand(10, 0, 1, 1, seed=1)
    y = dev(np.full(10, 7.0))
    bm.mul_(y, A, dev(np.zeros(0)), 1.0, 0.0)
    assert np.array_equal(y.cpu().numpy(), np.zeros(10))
    B = bm.brand(0, 10, 1, 1, seed=1)
    assert bm.matmul(B, dev(np.ones(10))).shape[0] == 0
    C = bm.brand(10, 10, 1, 1, seed=1)
    with pytest.raises(bm.DimensionMismatch):
        bm.mul_(dev(np.zeros(10)), C, dev(np.zeros(9)))
    with pytest.raises(bm.DimensionMismatch):
        bm.matmul(C, dev(np.zeros(11)))


def test_gbmv_large_properties(bm, oracle_ob):
    """Size-independent checks at a size the oracle cannot sweep exhaustively: linearity and a sampled
    exact comparison against OpenBLAS on a 2^22-row (4,3) matrix (the C2 configuration, scaled by 1/32)."""
    n, l, u = 1 << 22, 4, 3
    A = bm.brand(n, n, l, u, seed=7)
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
    y = torch.empty(n, dtype=torch.float64, device="cuda")
    bm.mul_(y, A, x)
    ref = np.zeros(n)
    gbmv(oracle_ob, "N", n, l, u, 1.0, A.banddata_host(), x.cpu().numpy(), 0.0, ref)
    assert np.array_equal(y.cpu().numpy(), ref)  # every one of the 4M rows, bit for bit
    y2 = torch.empty_like(y)
    bm.mul_(y2, A, 2.0 * x)  # scaling x by a power of two is exact
    assert torch.equal(y2, 2.0 * y)


def test_gbmv_host_entry(bm, oracle_c, rng):
    """bmb200_dgbmv_host (host arrays in, host array out; chunked upload) == oracle."""
    for (m, n, l, u, tr) in [(5000, 5000, 4, 3, "N"), (5000, 4000, 3, 9, "N"), (3000, 5000, 20, 20, "T"), (70000, 70000, 4, 3, "N")]:
        A = brand(rng, m, n, l, u, corners=np.nan)
        x = rng.standard_normal(n if tr == "N" else m)
        y0 = rng.standard_normal(m if tr == "N" else n)
        ref = y0.copy()
        gbmv(oracle_c, tr, m, l, u, 1.5, A.data, x, 0.5, ref)
        got = y0.copy()
        bm.gbmv_host(tr, m, l, u, 1.5, A.data, x, 0.5, got)
        if tr == "N":
            assert np.array_equal(got, ref)
        else:
            assert np.max(np.abs(got - ref) / np.maximum(1, np.abs(ref))) <= TOL
