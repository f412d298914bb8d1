"""GPU parity: lu / lu! / ldiv! / \\ through the C ABI vs the oracle.  Pivots AND factors bit-identical to
DGBTF2, solutions bit-identical to DGBTRS 'N'; transposed solves to tolerance; residual bound from
north_star (relative residual <= 1e-12 * cond).  Shapes from test/test_bandedlu.jl."""
import numpy as np
import pytest
import torch

import oracle
from oracle import Band, brand, ldiv, lu

from _util import golden_cases, kat_matrix

pytestmark = pytest.mark.gpu


def up(bm, Bd: Band):
    return bm.BandedMatrix.from_banddata(Bd.data, Bd.m, Bd.l, Bd.u)


def test_golden_lu(bm):
    for cid, c in golden_cases("lu"):
        n, l, u, nrhs, info = (int(v) for v in c["dims"])
        F = bm.lu(bm.BandedMatrix.from_banddata(c["data"], n, l, u))
        assert np.array_equal(F.ipiv, c["ipiv"]), cid                       # pivots bit-identical
        got = F.factors.banddata_host()
        if u > 64 and l >= 32:   # fixture produced by OpenBLAS' blocked DGBTRF: DGEMM rounding differs
            assert np.max(np.abs(got - c["ab"])) < 1e-10, cid
        else:
            assert np.array_equal(got, c["ab"]), cid                        # factors bit-identical
        Fg = bm.BandedLU(bm.BandedMatrix.from_banddata(c["ab"], n, l, l + u), c["ipiv"], 0)
        X = bm.to_colmajor(c["B"])
        bm.ldiv_(Fg, X)
        assert np.array_equal(X.cpu().numpy(), c["X"]), cid                 # solve bit-identical
        XT = bm.to_colmajor(c["B"])
        bm.ldiv_(Fg.T, XT)
        assert np.max(np.abs(XT.cpu().numpy() - c["XT"])) <= 1e-10 * np.max(np.abs(c["XT"])), cid


@pytest.mark.parametrize("shape", [(1000, 4, 3, 1), (10000, 4, 3, 1), (5000, 16, 16, 8), (3000, 5, 7, 5), (64, 3, 2, 2),
                                   (2000, 64, 64, 2), (1, 0, 0, 1), (5, 4, 4, 1), (2000, 0, 3, 2), (2000, 3, 0, 2),
                                   (4000, 1, 1, 4), (3000, 31, 0, 3), (3000, 33, 31, 2), (6, 7, 9, 1)])
def test_lu_solve_bit_identical(bm, oracle_c, rng, shape):
    n, l, u, nrhs = shape
    A = brand(rng, n, n, l, u)
    ab, ipiv, info = lu(oracle_c, A)
    F = bm.lu(up(bm, A))
    assert F.info == 0 and F.issuccess()
    assert (F.factors.l, F.factors.u) == (l, l + u)
    assert np.array_equal(F.ipiv, ipiv)
    assert np.array_equal(F.factors.banddata_host(), ab)
    B = np.asfortranarray(rng.standard_normal((n, nrhs)))
    ref = B.copy(order="F")
    ldiv(oracle_c, "N", ab, ipiv, l, u, ref)
    X = bm.to_colmajor(B)
    bm.ldiv_(F, X)
    # equal_nan: random triangular bands (l = 0 or u = 0) overflow identically in both implementations
    assert np.array_equal(X.cpu().numpy(), ref, equal_nan=True)
    refT = B.copy(order="F")
    ldiv(oracle_c, "T", ab, ipiv, l, u, refT)
    XT = bm.to_colmajor(B)
    bm.ldiv_(F.T, XT)
    if np.isfinite(refT).all():
        assert np.max(np.abs(XT.cpu().numpy() - refT)) <= 1e-9 * max(1e-300, np.max(np.abs(refT)))
    # vector right-hand side and A \ b (b not overwritten: test_bandedlu.jl:22-23)
    b = torch.as_tensor(B[:, 0].copy()).cuda()
    b_keep = b.clone()
    x = bm.solve(up(bm, A), b)
    assert torch.equal(b, b_keep)
    assert np.array_equal(x.cpu().numpy(), ref[:, 0], equal_nan=True)


def test_residual_bound(bm, rng):
    """north_star: relative residual <= 1e-12 * cond (here cond is computed densely at n = 400)."""
    n, l, u = 400, 16, 16
    A = brand(rng, n, n, l, u)
    D = A.dense()
    b = rng.standard_normal(n)
    x = bm.solve(up(bm, A), torch.as_tensor(b).cuda()).cpu().numpy()
    res = np.linalg.norm(D @ x - b) / (np.linalg.norm(D, 2) * np.linalg.norm(x))
    assert res <= 1e-12 * np.linalg.cond(D)
    assert res <= 1e-13


def test_lu_kat_pivot_vector(bm):
    """L,U,p = lu(A): the pivot vector equals dense LU's (test_bandedlu.jl:26-38) on the integer KAT matrix."""
    import scipy.linalg as sl

    D, v, X = kat_matrix()
    F = bm.lu(bm.BandedMatrix.from_dense(D, (2, 2)))
    _, piv = sl.lu_factor(D)
    assert np.array_equal(F.ipiv, piv + 1)
    sol = bm.solve(bm.BandedMatrix.from_dense(D, (2, 2)), bm.to_colmajor(X)).cpu().numpy()
    assert np.allclose(D @ sol, X, rtol=1e-10, atol=1e-9)


def test_singular_raises(bm):
    D = np.diag([1.0, 2.0, 0.0, 4.0, 5.0, 6.0])
    with pytest.raises(bm.LAPACKException) as e:
        bm.lu(bm.BandedMatrix.from_dense(D, (1, 1)))
    assert e.value.info == 3


def test_zero_size_and_nonsquare(bm):
    """test_bandedlu.jl:157-164 (0x0, 0x3, negative bandwidths) and :79-87 (non-square ⇒ DimensionMismatch)."""
    for (m, n, l, u) in [(0, 0, 1, 1), (0, 3, 1, 1), (0, 0, -1, -1)]:
        A = bm.BandedMatrix.zeros((m, n), (l, u))
        F = bm.lu(A)
        assert F.ipiv.size == 0 and not F.factors.data.any()
    A0 = bm.BandedMatrix.zeros((0, 0), (1, 1))
    assert bm.solve(A0, torch.zeros(0, dtype=torch.float64, device="cuda")).numel() == 0
    R = bm.brand(6, 7, 1, 1, seed=1)
    with pytest.raises(bm.DimensionMismatch):
        bm.solve(R, torch.zeros(6, dtype=torch.float64, device="cuda"))


def test_lu_large_c4_shape_scaled(bm, oracle_ob, rng):
    """C4's band (16,16) at n = 2^16 with 8 RHS: pivots / factors / solution bit-identical to OpenBLAS itself."""
    n, l, u, nrhs = 1 << 16, 16, 16, 8
    A = brand(rng, n, n, l, u)
    ab, ipiv, info = lu(oracle_ob, A)
    F = bm.lu(up(bm, A))
    assert np.array_equal(F.ipiv, ipiv)
    assert np.array_equal(F.factors.banddata_host(), ab)
    B = np.asfortranarray(rng.standard_normal((n, nrhs)))
    ref = B.copy(order="F")
    ldiv(oracle_ob, "N", ab, ipiv, l, u, ref)
    X = bm.to_colmajor(B)
    bm.ldiv_(F, X)
    assert np.array_equal(X.cpu().numpy(), ref)


@pytest.mark.parametrize("shape", [(1500, 150, 140, 3), (2000, 100, 65, 2), (1300, 40, 100, 2), (2500, 300, 200, 2), (900, 70, 70, 1)])
def test_wide_band_blocked_path_bit_identical_to_dgbtf2(bm, oracle_c, rng, shape):
    """Bands too wide for the shared-memory window: panel + trailing-update kernels, wide solve kernel.
    Per-element FMA order is DGBTF2's, so factors / pivots / solution equal the oracle bit for bit."""
    n, l, u, nrhs = shape
    A = brand(rng, n, n, l, u)
    ab, ipiv, info = lu(oracle_c, A)
    F = bm.lu(up(bm, A))
    assert np.array_equal(F.ipiv, ipiv)
    assert np.array_equal(F.factors.banddata_host(), ab)
    B = np.asfortranarray(rng.standard_normal((n, nrhs)))
    ref = B.copy(order="F")
    ldiv(oracle_c, "N", ab, ipiv, l, u, ref)
    X = bm.to_colmajor(B)
    bm.ldiv_(F, X)
    assert np.array_equal(X.cpu().numpy(), ref)


@pytest.mark.parametrize("shape", [(2600, 300, 200, 3), (5000, 1024, 1024, 2), (1800, 150, 140, 1), (4100, 129, 500, 2),
                                   (3001, 333, 217, 2), (1000, 500, 300, 1), (2000, 0, 400, 2), (2500, 400, 0, 2),
                                   (777, 130, 129, 1), (20011, 160, 150, 19),
                                   # mid-width interchange-free bands: beyond the register-window solves, now also on the cluster pipeline
                                   (3000, 100, 100, 4), (2500, 70, 90, 16), (1500, 128, 128, 2), (2200, 64, 10, 3), (1900, 20, 120, 5)])
def test_wide_band_dominant_optimistic_lu_and_blocked_solve(bm, oracle_c, rng, shape):
    """Diagonally dominant wide bands (the C5 regime): the pipelined factorisation takes its optimistic path
    (diagonal pivots, verified), the solve the panel-blocked interchange-free kernel.  Pivots = 1:n, factors and
    solution bit-identical to DGBTF2 / DGBTRS."""
    n, l, u, nrhs = shape
    A = brand(rng, n, n, l, u)
    A.data[u, :] += 2.0 * (l + u + 1)
    ab, ipiv, info = lu(oracle_c, A)
    assert np.array_equal(ipiv, np.arange(1, n + 1))
    F = bm.lu(up(bm, A))
    assert np.array_equal(F.ipiv, ipiv)
    assert np.array_equal(F.factors.banddata_host(), ab)
    B = np.asfortranarray(rng.standard_normal((n, nrhs)))
    ref = B.copy(order="F")
    ldiv(oracle_c, "N", ab, ipiv, l, u, ref)
    X = bm.to_colmajor(B)
    bm.ldiv_(F, X)
    assert np.array_equal(X.cpu().numpy(), ref)


def test_wide_band_violation_mid_factorisation_falls_back(bm, oracle_c, rng):
    """A dominant matrix with one weak diagonal far from the start: the optimistic path must notice, restore the panel
    and continue with searched pivots -- still DGBTF2's pivots and bits."""
    n, l, u = 3000, 200, 180
    A = brand(rng, n, n, l, u)
    A.data[u, :] += 2.0 * (l + u + 1)
    A.data[u, 1777] = 1e-3          # forces an interchange at column 1777
    ab, ipiv, info = lu(oracle_c, A)
    assert (ipiv != np.arange(1, n + 1)).any()
    F = bm.lu(up(bm, A))
    assert np.array_equal(F.ipiv, ipiv)
    assert np.array_equal(F.factors.banddata_host(), ab)


def laplacian_band(N):
    """examples/finitedifference_2d.jl:10-16,29: A = I - dt*(kron(D2,I)+kron(I,D2)), D2 = N^2*tridiag(1,-2,1), dt = 1/(4N^2):
    diagonal 2, bands +-1 and +-N equal -0.25 (0 across block edges). Returns the (2N+1) x N^2 band data."""
    n = N * N
    data = np.zeros((2 * N + 1, n), order="F")
    data[N, :] = 2.0
    j = np.arange(n)
    data[N - 1, 1:] = np.where(j[1:] % N != 0, -0.25, 0.0)        # superdiagonal: A[j-1, j]
    data[N + 1, :-1] = np.where((j[:-1] + 1) % N != 0, -0.25, 0.0)  # subdiagonal: A[j+1, j]
    data[0, N:] = -0.25
    data[2 * N, :-N] = -0.25
    return data


@pytest.mark.parametrize("N", [32, 64, 96])
def test_laplacian_lu_identity_pivots_and_residual(bm, oracle_ob, N):
    """Config C5 scaled down: strictly diagonally dominant => ipiv = 1:n; residual bound; vs OpenBLAS (blocked regime
    for N > 64: factors to tolerance, pivots exact)."""
    n = N * N
    data = laplacian_band(N)
    A = bm.BandedMatrix.from_banddata(data, n, N, N)
    F = bm.lu(A)
    assert np.array_equal(F.ipiv, np.arange(1, n + 1))
    ab, ipiv, info = lu(oracle_ob, Band(data, n, N, N))
    assert np.array_equal(F.ipiv, ipiv)
    assert np.max(np.abs(F.factors.banddata_host() - ab)) < 1e-12
    b = np.ones(n)
    x = bm.solve(A, torch.as_tensor(b).cuda()).cpu().numpy()
    r = b.copy()
    Ah = Band(data, n, N, N)
    oracle.gbmv(oracle_ob, "N", n, N, N, -1.0, data, x, 1.0, r)  # r = b - A x
    assert np.max(np.abs(r)) <= 1e-12 * np.max(np.abs(x)) * 3.0  # ||A||_inf = 3, cond ~ O(1)


@pytest.mark.parametrize("shape", [(5000, 16, 16, 64), (3000, 16, 16, 40), (4097, 4, 3, 33), (2000, 5, 7, 16), (700, 32, 32, 32),
                                   (1000, 31, 20, 17), (900, 0, 5, 32), (900, 6, 0, 32), (20, 4, 3, 32), (3, 2, 2, 16),
                                   (257, 1, 1, 100), (1500, 2, 1, 64), (800, 20, 40, 24)])
def test_solve_many_rhs_lane_kernel_bit_identical(bm, oracle_c, rng, shape):
    """nrhs >= 16 and a narrow band: the one-lane-per-RHS register-window kernels (gbtrs_lane.cu).  Partial RHS tiles,
    n smaller than the window / the prefetch distance, kl = 0 and ku = 0 included."""
    n, l, u, nrhs = shape
    A = brand(rng, n, n, l, u)
    ab, ipiv, info = lu(oracle_c, A)
    F = bm.lu(up(bm, A))
    assert np.array_equal(F.ipiv, ipiv)
    B = np.asfortranarray(rng.standard_normal((n, nrhs)))
    ref = B.copy(order="F")
    ldiv(oracle_c, "N", ab, ipiv, l, u, ref)
    X = bm.to_colmajor(B)
    bm.ldiv_(F, X)
    assert np.array_equal(X.cpu().numpy(), ref, equal_nan=True)


def test_solve_many_rhs_c4_scaled_vs_openblas(bm, oracle_ob, rng):
    """C4's band with 64 RHS at n = 2^15 against OpenBLAS dgbtrs itself."""
    n, l, u, nrhs = 1 << 15, 16, 16, 64
    A = brand(rng, n, n, l, u)
    ab, ipiv, info = lu(oracle_ob, A)
    F = bm.lu(up(bm, A))
    B = np.asfortranarray(rng.standard_normal((n, nrhs)))
    ref = B.copy(order="F")
    ldiv(oracle_ob, "N", ab, ipiv, l, u, ref)
    X = bm.to_colmajor(B)
    bm.ldiv_(F, X)
    assert np.array_equal(X.cpu().numpy(), ref)


@pytest.mark.parametrize("scale", [1e-300, 1e-160, 1e-30, 1e30, 1e150, 1e300])
def test_solve_extreme_scaling_division_paths(bm, oracle_c, rng, scale):
    """The backward sweep divides with a pre-refined reciprocal (divisor half hoisted off the dependency chain) and
    falls back to the stock operator outside the fast path's range: wildly scaled factors and right-hand sides
    (subnormal / huge quotients included) must still match DTBSV's true division bit for bit."""
    n, l, u, nrhs = 600, 3, 2, 4
    A = brand(rng, n, n, l, u)
    A.data[:] = A.data * scale
    ab, ipiv, info = lu(oracle_c, A)
    F = bm.lu(up(bm, A))
    assert np.array_equal(F.ipiv, ipiv)
    assert np.array_equal(F.factors.banddata_host(), ab)
    for bscale in (1.0, 1e-290, 1e290, 1e-308):
        B = np.asfortranarray(rng.standard_normal((n, nrhs)) * bscale)
        ref = B.copy(order="F")
        with np.errstate(all="ignore"):
            ldiv(oracle_c, "N", ab, ipiv, l, u, ref)
        X = bm.to_colmajor(B)
        bm.ldiv_(F, X)
        assert np.array_equal(X.cpu().numpy(), ref, equal_nan=True), (scale, bscale)


def test_blocked_solve_fast_division_is_the_ieee_quotient(bm, rng):
    """gbtrs_blocked.cu replaces the per-column division of the back substitution by two Markstein corrections of
    x * RN(1/d); the result must be the correctly rounded quotient bit for bit (DTBSV divides).
    Random operands over the whole exponent range, near-powers-of-two, all-ones significands, zeros, subnormals, Inf."""
    import ctypes as C

    hd = bm.handle(0)
    fn = hd.lib.bmb200_internal_divcheck
    n = 1 << 24
    m = rng.random(n) + 1.0
    e = rng.integers(-1000, 1000, n)
    x = np.ldexp(m, e) * rng.choice([-1.0, 1.0], n)
    d = np.ldexp(rng.random(n) + 1.0, rng.integers(-1000, 1000, n)) * rng.choice([-1.0, 1.0], n)
    # adversarial significands
    k = n // 16
    d[:k] = np.ldexp(2.0 - 2.0 ** -52, rng.integers(-600, 600, k))            # all ones
    d[k:2 * k] = np.ldexp(1.0 + 2.0 ** -52 * rng.integers(0, 4, k), rng.integers(-600, 600, k))
    x[2 * k:3 * k] = np.ldexp(2.0 - 2.0 ** -52 * rng.integers(1, 4, k), rng.integers(-600, 600, k))
    x[3 * k:3 * k + 8] = [0.0, -0.0, 5e-324, -1e-310, np.inf, -np.inf, np.nan, 1e-320]
    d[4 * k:4 * k + 6] = [5e-324, 1e-310, np.inf, 1e308, 1e-308, 3.0]
    # products that land near rounding boundaries: x = RN(q * d) for random q
    q = np.ldexp(rng.random(k) + 1.0, rng.integers(-100, 100, k))
    x[5 * k:6 * k] = q * d[5 * k:6 * k]
    dx, dd = torch.as_tensor(x).cuda(), torch.as_tensor(d).cuda()
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    hd.check(fn(hd.h, n, C.c_void_p(dx.data_ptr()), C.c_void_p(dd.data_ptr()), C.c_void_p(bad.data_ptr())), "divcheck")
    torch.cuda.synchronize()
    assert int(bad.item()) == 0
    # the slot-scheduled solve's division (gbtrs_slot.cu): q = fma(t, r_hi, t*r_lo) accepted only when one Markstein
    # correction reproduces it, IEEE division otherwise -- the accepted value must be the IEEE quotient on every input,
    # and nearly all ordinary operands must take the verified fast path
    bad2 = torch.zeros(2, dtype=torch.int64, device="cuda")
    hd.check(hd.lib.bmb200_internal_divcheck2(hd.h, n, C.c_void_p(dx.data_ptr()), C.c_void_p(dd.data_ptr()),
                                              C.c_void_p(bad2.data_ptr())), "divcheck2")
    torch.cuda.synchronize()
    assert int(bad2[0].item()) == 0
    assert int(bad2[1].item()) > 0.2 * n  # an operand beyond 2^+-500 (3/4 of this sample) takes the IEEE route


@pytest.mark.parametrize("PF,PB,W,RF,RB", [(2, 2, 1, 1, 1), (4, 4, 1, 1, 1), (4, 8, 2, 1, 1), (8, 8, 4, 1, 1), (4, 4, 2, 2, 2)])
@pytest.mark.parametrize("shape", [(1000, 16, 16, 5), (777, 4, 3, 9), (64, 3, 2, 2), (3000, 24, 8, 3), (130, 5, 7, 17),
                                   (1, 0, 0, 1), (5, 4, 4, 1), (2000, 0, 3, 2), (2000, 3, 0, 2), (4097, 1, 1, 4),
                                   (63, 2, 30, 1), (65, 20, 12, 33)])
def test_slot_scheduled_solve_bit_identical(bm, oracle_c, rng, shape, PF, PB, W, RF, RB):
    """gbtrs_slot.cu, (PF, PB, W, RF, RB) variants through the internal hook: solutions bit-identical to DGBTRS 'N' with
    interchanges, ragged n around the 64-column schedule stages, kl = 0, ku = 0, more RHS than a CTA holds."""
    import ctypes as C

    n, l, u, nrhs = shape
    if l + PF > 32:
        pytest.skip("window does not fit a warp for this P")
    A = brand(rng, n, n, l, u)
    ab, ipiv, info = lu(oracle_c, A)
    B = np.asfortranarray(rng.standard_normal((n, nrhs)))
    B[rng.integers(0, n, 3), 0] = 0.0  # exact zeros take the verified-division fallback
    ref = B.copy(order="F")
    ldiv(oracle_c, "N", ab, ipiv, l, u, ref)
    hd = bm.handle(0)
    dab = torch.as_tensor(np.ascontiguousarray(ab.T)).cuda()  # (n, ldab): column j of AB contiguous
    dip = torch.as_tensor(ipiv.astype(np.int64)).cuda()
    X = bm.to_colmajor(B)
    rc = hd.lib.bmb200_internal_gbtrs_slot(hd.h, PF, PB, W, RF, RB, n, l, u, nrhs, C.c_void_p(dab.data_ptr()), ab.shape[0],
                                           C.c_void_p(dip.data_ptr()), C.c_void_p(X.data_ptr()), max(1, n))
    hd.check(rc, "internal_gbtrs_slot")
    torch.cuda.synchronize()
    assert np.array_equal(X.cpu().numpy(), ref, equal_nan=True)


def test_slot_solve_sparse_rhs_and_nonfinite(bm, oracle_c, rng):
    """Unit-vector right-hand sides (long runs of exact zeros: the fast division's zero route), and an Inf / NaN planted
    in b: both must propagate exactly as in DGBTRS (absent updates never touch a row before it enters the window)."""
    n, l, u = 500, 6, 5
    A = brand(rng, n, n, l, u)
    ab, ipiv, info = lu(oracle_c, A)
    B = np.zeros((n, 6), order="F")
    B[0, 0] = 1.0
    B[n - 1, 1] = -2.0
    B[250, 2] = 3.0
    B[:, 3] = rng.standard_normal(n)
    B[300, 3] = np.inf
    B[:, 4] = rng.standard_normal(n)
    B[100, 4] = np.nan
    B[:, 5] = -0.0
    ref = B.copy(order="F")
    ldiv(oracle_c, "N", ab, ipiv, l, u, ref)
    F = bm.BandedLU(bm.BandedMatrix.from_banddata(ab, n, l, l + u), ipiv, 0)
    X = bm.to_colmajor(B)
    bm.ldiv_(F, X)
    got = X.cpu().numpy()
    assert np.array_equal(got, ref, equal_nan=True)
    assert np.array_equal(np.signbit(got[:, 5]), np.signbit(ref[:, 5]))  # signed zeros survive too


# ---- round 2: strip-resident wide-band LU (gbtrf_strip.cu) and multi-warp narrow LU (gbtrf_mw.cu) ----
@pytest.mark.parametrize("shape", [(3000, 200, 180), (2100, 64, 257), (4000, 513, 48)])
def test_strip_lu_in_place_entry_uses_device_copy(bm, oracle_c, rng, shape):
    """lu!(A) on an already widened matrix (bmb200_dgbtrf without the source): the optimistic strip kernel keeps a
    device-side copy of the band; pivots = 1:n and factors bit-identical to DGBTF2.  Then the same with a weak diagonal:
    the copy is restored and the general path gives DGBTF2's pivots and bits."""
    n, l, u = shape
    A = brand(rng, n, n, l, u)
    A.data[u, :] += 2.0 * (l + u + 1)
    for weak in (None, n // 2 + 3):
        if weak is not None:
            A.data[u, weak] = 1e-3
        ab, ipiv, info = lu(oracle_c, A)
        W = bm.BandedMatrix.zeros((n, n), (l, l + u))
        W.data[:, l:] = torch.as_tensor(np.ascontiguousarray(A.data.T)).cuda()
        F = bm.lu_(W)
        assert np.array_equal(F.ipiv, ipiv), weak
        assert np.array_equal(F.factors.banddata_host(), ab), weak
        assert (weak is None) == bool((ipiv == np.arange(1, n + 1)).all())


def test_wide_band_nan_entry_no_cuda_error(bm, rng):
    """A NaN inside a dominant wide band: the strip kernel treats it as a violation, the general path factors on; the
    call must return (NaN in the factors), the pivots must stay in range, and the context must stay usable."""
    n, l, u = 2500, 130, 120
    A = brand(rng, n, n, l, u)
    A.data[u, :] += 2.0 * (l + u + 1)
    A.data[u + 7, 900] = np.nan
    F = bm.lu(up(bm, A))
    assert F.ipiv.min() >= 1 and F.ipiv.max() <= n
    assert np.isnan(F.factors.banddata_host()).any()
    B = brand(rng, 500, 500, 4, 3)          # the handle still works
    ab, ipiv, info = lu(oracle.backend("C"), B)
    F2 = bm.lu(up(bm, B))
    assert np.array_equal(F2.ipiv, ipiv) and np.array_equal(F2.factors.banddata_host(), ab)


@pytest.mark.parametrize("shape", [(900, 1300, 16, 16), (1300, 900, 16, 16), (40, 40, 16, 16), (17, 64, 5, 2), (3000, 3000, 31, 1),
                                   (3000, 3000, 1, 31), (2000, 2000, 2, 1), (515, 515, 9, 6), (8, 8, 4, 3)])
def test_multi_warp_narrow_lu_matches_single_warp_and_oracle(bm, oracle_c, rng, shape):
    """gbtrf_mw.cu (chain warp + far warps) against the oracle and against the single-warp kernel it replaces, rectangular and
    tiny shapes included (entering rows that do not exist, U rows cut by the matrix edge)."""
    m, n, l, u = shape
    A = brand(rng, m, n, l, u)
    # rectangular: straight to the oracle's DGBTF2 (its lu() wrapper mirrors the reference's checksquare of `\`)
    ab = np.zeros((2 * l + u + 1, n), order="F")
    ab[l:, :] = A.data
    ipiv = np.zeros(min(m, n), dtype=np.int64)
    oracle_c.gbtrf(m, n, l, u, ab, ab.shape[0], ipiv)
    hd = bm.handle(0)
    F = bm.lu(up(bm, A))
    assert np.array_equal(F.ipiv, ipiv)
    got, inm = F.factors.banddata_host(), np.zeros(ab.shape, dtype=bool)
    rr, jj = np.meshgrid(np.arange(ab.shape[0]), np.arange(n), indexing="ij")
    inm = (jj + rr - (l + u) >= 0) & (jj + rr - (l + u) < m)     # slots outside the matrix are never touched by either side
    assert np.array_equal(got[inm], ab[inm])
    hd.tune("gbtrf_nomw", 1)
    try:
        F1 = bm.lu(up(bm, A))
    finally:
        hd.tune("reset", 0)
    assert np.array_equal(F1.ipiv, F.ipiv)
    assert np.array_equal(F1.factors.banddata_host(), F.factors.banddata_host())
