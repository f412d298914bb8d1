/*
 * bmoracle.c -- CPU ORACLE for the banded hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing under oracle/ is part of the product.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may build, load or call it, and
 * there only as the checker.  The shipped path is bandedmatrices.jl_b200/csrc (CUDA).
 *
 * What is restated here.  BandedMatrices.jl contains no arithmetic for this path: its
 * drivers (src/generic/matmul.jl, src/banded/gbmm.jl, src/banded/BandedLU.jl,
 * src/banded/linalg.jl) ccall Fortran BLAS/LAPACK entry points of OpenBLAS
 * (third-party, not under /root/reference; version unpinned by the reference - Julia 1.10
 * bundles OpenBLAS 0.3.23, current 1.x 0.3.29).  This file restates
 *   - the published reference-BLAS/LAPACK algorithms behind those entry points
 *     (DGBMV, DGBTF2/DGBTRF, DGBTRS, DTBSV) with the operation order OpenBLAS 0.3.30
 *     was observed to use (FMA placement, reciprocal-vs-division, first-max pivot), and
 *   - the reference's own Julia driver loops that sit above them
 *     (_gbmm! four column regimes, _lu widening copy, _fill_lmul! beta rule).
 * Pinning: tests/test_oracle_pin.py checks every routine here bit-for-bit (gbmv 'N',
 * gbtf2, gbtrs 'N') or to 1e-13 (transposed variants, blocked regime) against the
 * OpenBLAS 0.3.30 ILP64 library shipped inside numpy in this image (same Fortran
 * symbols Julia calls), and against the reference's deterministic integer-valued
 * known-answer test (test/test_linalg.jl:51-97).  Each function cites the reference
 * call site it serves.
 *
 * Storage conventions (0-based here):
 *   band storage   A[k,j] at a[(ku + k - j) + j*lda]            (src/banded/BandedMatrix.jl:414-419)
 *   LU storage     A[k,j] at ab[(kl + ku + k - j) + j*ldab], ldab >= 2kl+ku+1
 *   ipiv           1-based, ipiv[j] in j+1 .. j+1+kl             (LAPACK convention, BandedLU.jl:12)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t i64;

static inline i64 imin(i64 a, i64 b) { return a < b ? a : b; }
static inline i64 imax(i64 a, i64 b) { return a > b ? a : b; }

/* y <- alpha*op(A)*x + beta*y.
 * Serves BLAS.gbmv! at src/generic/matmul.jl:21-23 and the pointer wrapper src/blas.jl:16-28.
 * 'N': beta==0 overwrites (NaN in y does not propagate); then one AXPY per column with
 * t = alpha*x[j] rounded and y[i] = fma(t, A[i,j], y[i])  (OpenBLAS dgbmv_n + daxpy FMA kernel).
 * 'T': reference-BLAS dot per column (OpenBLAS uses a SIMD dot whose order is unspecified,
 * so 'T' is compared to tolerance, not bits).
 * Out-of-matrix corner slots of band storage are never read. */
int oracle_dgbmv(char trans, i64 m, i64 n, i64 kl, i64 ku, double alpha,
                 const double *a, i64 lda, const double *x, i64 incx,
                 double beta, double *y, i64 incy)
{
    int tr = (trans == 'T' || trans == 't' || trans == 'C' || trans == 'c');
    if (!tr && !(trans == 'N' || trans == 'n')) return -1;
    if (m < 0) return -2;
    if (n < 0) return -3;
    if (kl < 0) return -4;
    if (ku < 0) return -5;
    if (lda < kl + ku + 1) return -8;
    if (incx == 0) return -10;
    if (incy == 0) return -13;
    if (m == 0 || n == 0) return 0;
    i64 lenx = tr ? m : n, leny = tr ? n : m;
    const double *xx = incx > 0 ? x : x - (lenx - 1) * incx;
    double *yy = incy > 0 ? y : y - (leny - 1) * incy;
    if (beta != 1.0) {
        for (i64 i = 0; i < leny; ++i)
            yy[i * incy] = (beta == 0.0) ? 0.0 : beta * yy[i * incy];
    }
    if (alpha == 0.0) return 0;
    if (!tr) {
        i64 jend = imin(n, m + ku);
        for (i64 j = 0; j < jend; ++j) {
            double t = alpha * xx[j * incx];
            i64 i0 = imax(0, j - ku), i1 = imin(m - 1, j + kl);
            const double *col = a + j * lda + (ku - j);
            for (i64 i = i0; i <= i1; ++i)
                yy[i * incy] = fma(t, col[i], yy[i * incy]);
        }
    } else {
        for (i64 j = 0; j < n; ++j) {
            i64 i0 = imax(0, j - ku), i1 = imin(m - 1, j + kl);
            const double *col = a + j * lda + (ku - j);
            double temp = 0.0;
            for (i64 i = i0; i <= i1; ++i)
                temp = fma(col[i], xx[i * incx], temp);
            yy[j * incy] = fma(alpha, temp, yy[j * incy]);
        }
    }
    return 0;
}

/* _fill_lmul!(beta, view) : src/generic/utils.jl:29 -- beta==0 zero-fills, else scales. */
void oracle_fill_lmul(double beta, double *c, i64 rows, i64 cols, i64 ldc)
{
    for (i64 j = 0; j < cols; ++j)
        for (i64 i = 0; i < rows; ++i)
            c[i + j * ldc] = (beta == 0.0) ? 0.0 : beta * c[i + j * ldc];
}

/* _gbmm!(alpha, A_data, B_data, beta, C_data, (n,nu,m), (Al,Au), (Bl,Bu), (Cl,Cu))
 * src/banded/gbmm.jl:296-340: one gbmv per column of C in three regimes, then the
 * trailing beta-scale.  A is n x nu, B is nu x m, C is n x m; all band widths >= 0 and
 * (Cl,Cu) = min((n-1,m-1),(Al+Bl,Au+Bu)) is guaranteed by the gbmm! driver (gbmm.jl:229).
 * j below is the reference's 1-based column index. */
int oracle_gbmm(i64 n, i64 nu, i64 m, i64 Al, i64 Au, i64 Bl, i64 Bu, i64 Cl, i64 Cu,
                double alpha, const double *a, i64 sta, const double *b, i64 stb,
                double beta, double *c, i64 stc)
{
    i64 j;
    /* A11_Btop_Ctop_gbmv!  gbmm.jl:18-45, loop :306-313 */
    for (j = 1; j <= imin(m, 1 + Bu); ++j)
        oracle_dgbmv('N', imin(Cl + j, n), imin(Bl + j, nu), Al, Au, alpha, a, sta,
                     b + ((j - 1) * stb + Bu - j + 1), 1, beta,
                     c + ((j - 1) * stc + Cu - j + 1), 1);
    /* Atop_Bmid_Ctop_gbmv!  gbmm.jl:57-86, loop :318-325 */
    for (j = 2 + Bu; j <= imin(imin(1 + Cu, nu + Bu), m); ++j)
        oracle_dgbmv('N', imin(Cl + j, n), imin(Bl + Bu + 1, nu - j + Bu + 1),
                     Al + j - Bu - 1, Au - j + Bu + 1, alpha,
                     a + (j - Bu - 1) * sta, sta, b + (j - 1) * stb, 1, beta,
                     c + ((j - 1) * stc + Cu - j + 1), 1);
    /* Amid_Bmid_Cmid_gbmv!  gbmm.jl:99-127, loop :329-336 */
    for (j = 2 + Cu; j <= imin(imin(m, nu + Bu), n + Cu); ++j) {
        i64 p = j - Bu;
        oracle_dgbmv('N', imin(Cl + Cu + 1, n - j + Cu + 1), imin(Bl + Bu + 1, nu - p + 1),
                     Al + Au, 0, alpha, a + (j - Bu - 1) * sta, sta,
                     b + (j - 1) * stb, 1, beta, c + (j - 1) * stc, 1);
    }
    /* _fill_lmul!(beta, view(C_data, :, nu+Bu+1:min(m,n+Cu)))  gbmm.jl:339 */
    i64 j0 = nu + Bu + 1, j1 = imin(m, n + Cu);
    if (j1 >= j0)
        oracle_fill_lmul(beta, c + (j0 - 1) * stc, Cl + Cu + 1, j1 - j0 + 1, stc);
    return 0;
}

/* BandedMatrix{T}(A,(l,l+u)) widening copy used by _lu: src/banded/BandedLU.jl:108-111,
 * ctor src/banded/BandedMatrix.jl:222-232.  Source (l+u+1) x n band data -> (2l+u+1) x n with
 * the original rows in rows l..2l+u and the top l rows zero. */
void oracle_band_widen(i64 n, i64 l, i64 u, const double *a, i64 lda, double *ab, i64 ldab)
{
    for (i64 j = 0; j < n; ++j) {
        for (i64 r = 0; r < l; ++r) ab[r + j * ldab] = 0.0;
        for (i64 r = 0; r < l + u + 1; ++r) ab[l + r + j * ldab] = a[r + j * lda];
    }
}

/* DGBTF2 -- unblocked partial-pivot band LU (reference LAPACK, compiled into OpenBLAS).
 * Serves LAPACK.gbtrf! at src/banded/BandedLU.jl:98 whenever ILAENV gives NB<=1 or NB>kl,
 * i.e. ku <= 64 or kl < 32 (configs C1, C4).  1-based indices in comments follow SURVEY.md A.3.
 * IDAMAX = FIRST maximum of |.|; multipliers scaled by the reciprocal; DGER update is
 * fma(-u, l, a) per element, columns ascending. */
int oracle_dgbtf2(i64 m, i64 n, i64 kl, i64 ku, double *ab, i64 ldab, i64 *ipiv)
{
    i64 kv = ku + kl;
    if (m < 0) return -1;
    if (n < 0) return -2;
    if (kl < 0) return -3;
    if (ku < 0) return -4;
    if (ldab < kl + kv + 1) return -6;
    if (m == 0 || n == 0) return 0;
#define AB(i, j) ab[((i) - 1) + ((j) - 1) * ldab] /* 1-based */
    int info = 0;
    for (i64 j = ku + 2; j <= imin(kv, n); ++j)
        for (i64 i = kv - j + 2; i <= kl; ++i) AB(i, j) = 0.0;
    i64 ju = 1;
    for (i64 j = 1; j <= imin(m, n); ++j) {
        if (j + kv <= n)
            for (i64 i = 1; i <= kl; ++i) AB(i, j + kv) = 0.0;
        i64 km = imin(kl, m - j);
        i64 jp = 1;
        double best = fabs(AB(kv + 1, j));
        for (i64 i = 2; i <= km + 1; ++i) {
            double v = fabs(AB(kv + i, j));
            if (v > best) { best = v; jp = i; }   /* NaN never wins, like IDAMAX */
        }
        ipiv[j - 1] = jp + j - 1;
        if (AB(kv + jp, j) != 0.0) {
            ju = imax(ju, imin(j + ku + jp - 1, n));
            if (jp != 1)
                for (i64 c = 0; c <= ju - j; ++c) { /* DSWAP, stride ldab-1 */
                    double t = AB(kv + jp - c, j + c);
                    AB(kv + jp - c, j + c) = AB(kv + 1 - c, j + c);
                    AB(kv + 1 - c, j + c) = t;
                }
            if (km > 0) {
                double r = 1.0 / AB(kv + 1, j);
                for (i64 i = 1; i <= km; ++i) AB(kv + 1 + i, j) *= r;
                for (i64 c = 1; c <= ju - j; ++c) {
                    double t = -AB(kv + 1 - c, j + c);
                    for (i64 i = 1; i <= km; ++i) /* OpenBLAS dger = one FMA axpy per column */
                        AB(kv + 1 - c + i, j + c) = fma(t, AB(kv + 1 + i, j), AB(kv + 1 - c + i, j + c));
                }
            }
        } else if (info == 0) {
            info = (int)j;
        }
    }
#undef AB
    return info;
}

/* DGBTRF entry point (BandedLU.jl:98).  The wide-band regime (ku > 64 and kl >= 32; config C5)
 * uses LAPACK's blocked algorithm in OpenBLAS (NB=32; DTRSM+DGEMM), whose factors differ from the
 * unblocked ones only by rounding and whose pivots are identical unless two candidates tie to
 * rounding.  The restatement keeps ONE arithmetic definition (DGBTF2 order) for both regimes;
 * the pin test compares it bit-for-bit in the unblocked regime and pivots-exact + 1e-12 in the
 * blocked one. */
int oracle_dgbtrf(i64 m, i64 n, i64 kl, i64 ku, double *ab, i64 ldab, i64 *ipiv)
{
    return oracle_dgbtf2(m, n, kl, ku, ab, ldab, ipiv);
}

/* DGBTRS -- solve with the band LU factors.  Serves LAPACK.gbtrs! at
 * src/banded/linalg.jl:28 ('N'), :46 ('T'), :62 ('C' == 'T' for real).
 * 'N' forward: swap then DGER (fma(-b_j, l, b)); backward: DTBSV upper/no-trans/non-unit with
 * TRUE division by the diagonal.  'T': DTBSV upper/trans then the L^T sweep with DGEMV-T dots
 * (order unspecified in OpenBLAS => tolerance). */
int oracle_dgbtrs(char trans, i64 n, i64 kl, i64 ku, i64 nrhs, const double *ab, i64 ldab,
                  const i64 *ipiv, double *b, i64 ldb)
{
    int tr = (trans == 'T' || trans == 't' || trans == 'C' || trans == 'c');
    if (!tr && !(trans == 'N' || trans == 'n')) return -1;
    if (n < 0) return -2;
    if (kl < 0) return -3;
    if (ku < 0) return -4;
    if (nrhs < 0) return -5;
    if (ldab < 2 * kl + ku + 1) return -7;
    if (ldb < imax(1, n)) return -10;
    if (n == 0 || nrhs == 0) return 0;
    i64 kv = kl + ku, kd = kv + 1; /* kd = 1-based row of the diagonal */
#define AB(i, j) ab[((i) - 1) + ((j) - 1) * ldab]
#define B(i, c) b[((i) - 1) + ((c) - 1) * ldb]
    if (!tr) {
        if (kl > 0)
            for (i64 j = 1; j <= n - 1; ++j) {
                i64 lm = imin(kl, n - j), p = ipiv[j - 1];
                if (p != j)
                    for (i64 c = 1; c <= nrhs; ++c) { double t = B(p, c); B(p, c) = B(j, c); B(j, c) = t; }
                for (i64 c = 1; c <= nrhs; ++c) {
                    double t = -B(j, c);
                    for (i64 i = 1; i <= lm; ++i) B(j + i, c) = fma(t, AB(kd + i, j), B(j + i, c));
                }
            }
        for (i64 c = 1; c <= nrhs; ++c)
            for (i64 j = n; j >= 1; --j) {
                B(j, c) = B(j, c) / AB(kv + 1, j); /* OpenBLAS dtbsv_NUN: divide, then FMA axpy */
                double t = -B(j, c);
                for (i64 i = j - 1; i >= imax(1, j - kv); --i)
                    B(i, c) = fma(t, AB(kv + 1 + i - j, j), B(i, c));
            }
    } else {
        for (i64 c = 1; c <= nrhs; ++c)
            for (i64 j = 1; j <= n; ++j) {
                double temp = B(j, c);
                for (i64 i = imax(1, j - kv); i <= j - 1; ++i)
                    temp = fma(-AB(kv + 1 + i - j, j), B(i, c), temp);
                B(j, c) = temp / AB(kv + 1, j);
            }
        if (kl > 0)
            for (i64 j = n - 1; j >= 1; --j) {
                i64 lm = imin(kl, n - j), p = ipiv[j - 1];
                for (i64 c = 1; c <= nrhs; ++c) {
                    double temp = 0.0;
                    for (i64 i = 1; i <= lm; ++i) temp = fma(AB(kd + i, j), B(j + i, c), temp);
                    B(j, c) = B(j, c) - temp;
                }
                if (p != j)
                    for (i64 c = 1; c <= nrhs; ++c) { double t = B(p, c); B(p, c) = B(j, c); B(j, c) = t; }
            }
    }
#undef AB
#undef B
    return 0;
}

/* x <- inv(T)*x, T = banded triangle in BLAS triangular-band storage (tbsv!, src/blas.jl:109-141; reached from
 * ldiv!(UpperTriangular/LowerTriangular{BandedMatrix}, x), src/tribanded.jl:75-96 ('T' for row-major layouts), and from the
 * back substitution inside dgbtrs).  Storage (0-based): 'U' T[i,j] at a[(k + i - j) + j*lda], 'L' T[i,j] at a[(i - j) + j*lda].
 * OpenBLAS driver/level2/tbsv_{U,L}.c: per column, true division by the diagonal (unless unit), then one FMA axpy. */
int oracle_dtbsv(char uplo, char trans, char diag, i64 n, i64 k, const double *a, i64 lda, double *x)
{
    const int up = (uplo == 'U' || uplo == 'u'), unit = (diag == 'U' || diag == 'u');
    const int tr = (trans == 'T' || trans == 't' || trans == 'C' || trans == 'c');
    if (!up && !(uplo == 'L' || uplo == 'l')) return -1;
    if (!tr && !(trans == 'N' || trans == 'n')) return -2;
    if (!unit && !(diag == 'N' || diag == 'n')) return -3;
    if (n < 0) return -4;
    if (k < 0) return -5;
    if (lda < k + 1) return -7;
    if (tr) {
        /* reference-BLAS DTBSV 'T': one dot product per column (OpenBLAS uses a SIMD dot whose order is unspecified, so
         * 'T' is compared to tolerance, not bits): x_j = (x_j - sum_i T[i,j] x_i) / T[j,j] */
        if (up) {
            for (i64 j = 0; j < n; ++j) {
                double acc = 0.0;
                for (i64 i = imax(0, j - k); i < j; ++i) acc = fma(a[(k + i - j) + j * lda], x[i], acc);
                x[j] = x[j] - acc;
                if (!unit) x[j] = x[j] / a[k + j * lda];
            }
        } else {
            for (i64 j = n - 1; j >= 0; --j) {
                double acc = 0.0;
                for (i64 i = imin(n - 1, j + k); i > j; --i) acc = fma(a[(i - j) + j * lda], x[i], acc);
                x[j] = x[j] - acc;
                if (!unit) x[j] = x[j] / a[j * lda];
            }
        }
        return 0;
    }
    if (up) {
        for (i64 j = n - 1; j >= 0; --j) {
            if (!unit) x[j] = x[j] / a[k + j * lda];
            const double t = -x[j];
            for (i64 i = j - 1; i >= imax(0, j - k); --i) x[i] = fma(t, a[(k + i - j) + j * lda], x[i]);
        }
    } else {
        for (i64 j = 0; j < n; ++j) {
            if (!unit) x[j] = x[j] / a[j * lda];
            const double t = -x[j];
            for (i64 i = j + 1; i <= imin(n - 1, j + k); ++i) x[i] = fma(t, a[(i - j) + j * lda], x[i]);
        }
    }
    return 0;
}

/* x <- T*x (tbmv!, src/blas.jl:71-105; reached from lmul!(UpperTriangular/LowerTriangular{BandedMatrix}, x),
 * src/tribanded.jl:47-55).  OpenBLAS driver/level2/tbmv_{U,L}.c ('N') sweeps the columns (ascending for 'U',
 * descending for 'L'): x[j] is first axpy'd into the rows it reaches and then scaled by the diagonal, so row i ends as
 * d_i*x_i (one rounded product; x_i itself if unit) followed by fma(x_j, T[i,j], .) over j = i+1, i+2, ... ('U') or
 * j = i-1, i-2, ... ('L') with the ORIGINAL x_j. */
int oracle_dtbmv(char uplo, char trans, char diag, i64 n, i64 k, const double *a, i64 lda, double *x)
{
    const int up = (uplo == 'U' || uplo == 'u'), unit = (diag == 'U' || diag == 'u');
    const int tr = (trans == 'T' || trans == 't' || trans == 'C' || trans == 'c');
    if (!up && !(uplo == 'L' || uplo == 'l')) return -1;
    if (!tr && !(trans == 'N' || trans == 'n')) return -2;
    if (!unit && !(diag == 'N' || diag == 'n')) return -3;
    if (n < 0) return -4;
    if (k < 0) return -5;
    if (lda < k + 1) return -7;
    if (n == 0) return 0;
    double *y = (double *)malloc((size_t)n * sizeof(double));
    if (!y) return -100;
    for (i64 i = 0; i < n; ++i) {
        double acc;
        if (tr) { /* (T^T x)_i = d_i x_i + sum over column i of T (dot; tolerance, see oracle_dtbsv) */
            double dot = 0.0;
            if (up) for (i64 r = imax(0, i - k); r < i; ++r) dot = fma(a[(k + r - i) + i * lda], x[r], dot);
            else for (i64 r = i + 1; r <= imin(n - 1, i + k); ++r) dot = fma(a[(r - i) + i * lda], x[r], dot);
            acc = (unit ? x[i] : x[i] * a[(up ? k : 0) + i * lda]) + dot;
        } else if (up) {
            acc = unit ? x[i] : x[i] * a[k + i * lda];
            for (i64 j = i + 1; j <= imin(n - 1, i + k); ++j) acc = fma(x[j], a[(k + i - j) + j * lda], acc);
        } else {
            acc = unit ? x[i] : x[i] * a[i * lda];
            for (i64 j = i - 1; j >= imax(0, i - k); --j) acc = fma(x[j], a[(i - j) + j * lda], acc);
        }
        y[i] = acc;
    }
    memcpy(x, y, (size_t)n * sizeof(double));
    free(y);
    return 0;
}

/* y <- alpha*S*x + beta*y, S symmetric, `uplo` triangle in triangular-band storage (sbmv!, src/blas.jl:36-66; mul! of
 * Symmetric{<:BandedMatrix}, src/symbanded/symbanded.jl:72-93).  Reference-BLAS DSBMV loop order; OpenBLAS pairs an axpy with a
 * SIMD dot per column (order unspecified), so the pin is to tolerance. */
int oracle_dsbmv(char uplo, i64 n, i64 k, double alpha, const double *a, i64 lda, const double *x, double beta, double *y)
{
    const int up = (uplo == 'U' || uplo == 'u');
    if (!up && !(uplo == 'L' || uplo == 'l')) return -1;
    if (n < 0) return -2;
    if (k < 0) return -3;
    if (lda < k + 1) return -6;
    for (i64 i = 0; i < n; ++i) y[i] = (beta == 0.0) ? 0.0 : beta * y[i];
    if (alpha == 0.0) return 0;
    for (i64 j = 0; j < n; ++j) {
        const double t1 = alpha * x[j];
        double t2 = 0.0;
        if (up) {
            for (i64 i = imax(0, j - k); i < j; ++i) {
                y[i] = fma(t1, a[(k + i - j) + j * lda], y[i]);
                t2 = fma(a[(k + i - j) + j * lda], x[i], t2);
            }
            y[j] = y[j] + t1 * a[k + j * lda] + alpha * t2;
        } else {
            y[j] = y[j] + t1 * a[j * lda];
            for (i64 i = j + 1; i <= imin(n - 1, j + k); ++i) {
                y[i] = fma(t1, a[(i - j) + j * lda], y[i]);
                t2 = fma(a[(i - j) + j * lda], x[i], t2);
            }
            y[j] = y[j] + alpha * t2;
        }
    }
    return 0;
}

/* banded_mul! triple loop (src/generic/matmul.jl:143-172): the semantic definition of
 * banded x banded used as a second, independent check of oracle_gbmm.  C gets zeros in bands
 * beyond (Al+Bl, Au+Bu).  Band widths may exceed the matrix size; all must be >= 0 here. */
void oracle_banded_mul(i64 Am, i64 An, i64 Bn, i64 Al, i64 Au, i64 Bl, i64 Bu, i64 Cl_, i64 Cu_,
                       const double *a, i64 lda, const double *b, i64 ldb, double *c, i64 ldc)
{
    i64 Cl = Al + Bl, Cu = Au + Bu;
    for (i64 j = 0; j < Bn; ++j)
        for (i64 k = imax(0, j - Cu_); k <= imin(Am - 1, j + Cl_); ++k) {
            double tmp = 0.0;
            if (k >= j - Cu && k <= j + Cl) {
                i64 v0 = imax(imax(0, k - Al), j - Bu), v1 = imin(imin(An - 1, k + Au), j + Bl);
                for (i64 v = v0; v <= v1; ++v)
                    tmp = tmp + a[(Au + k - v) + v * lda] * b[(Bu + v - j) + j * ldb];
            }
            c[(Cu_ + k - j) + j * ldc] = tmp;
        }
}


/* ---------------------------------------------------------------------------------------------------------------------
 * Banded Cholesky: pbtrf! / pbtrs! (src/lapack.jl:268-332) behind cholesky(Symmetric(::BandedMatrix)) and its ldiv!
 * (src/symbanded/BandedCholesky.jl:2-13, 72-80).  Reference-LAPACK DPBTF2 (the unblocked algorithm DPBTRF runs for kd < 32; its
 * blocked form for wider bands differs by DGEMM / DSYRK rounding only):
 *   for j: ajj = A[j,j]; ajj <= 0 -> info = j+1, stop; d = sqrt(ajj); row j of U (resp. column j of L) *= 1/d (DSCAL by the
 *   reciprocal); trailing kn x kn triangle -= x x^T (DSYR, OpenBLAS: one AXPY per column with t = -x[k] and a FMA per entry).
 * uplo 'U': A[i,k] (i <= k) at ab[(kd + i - k) + k*ldab];  'L': A[i,k] (i >= k) at ab[(i - k) + k*ldab].
 * --------------------------------------------------------------------------------------------------------------------- */
int oracle_dpbtf2(char uplo, i64 n, i64 kd, double *ab, i64 ldab)
{
    const int up = (uplo == 'U' || uplo == 'u');
    if (!up && !(uplo == 'L' || uplo == 'l')) return -1;
    if (n < 0) return -2;
    if (kd < 0) return -3;
    if (ldab < kd + 1) return -5;
#define PB(i, k) ab[up ? ((kd + (i) - (k)) + (k) * ldab) : (((k) - (i)) + (i) * ldab)] /* symmetric entry (i <= k) */
    for (i64 j = 0; j < n; ++j) {
        double ajj = PB(j, j);
        if (ajj <= 0.0) return (int)(j + 1);
        ajj = sqrt(ajj);
        PB(j, j) = ajj;
        const i64 kn = imin(kd, n - 1 - j);
        const double rinv = 1.0 / ajj;
        for (i64 c = 1; c <= kn; ++c) PB(j, j + c) = PB(j, j + c) * rinv;
        for (i64 c = 1; c <= kn; ++c) {          /* DSYR: column (resp. row) j+c of the trailing triangle */
            const double xc = PB(j, j + c);
            if (xc == 0.0) continue;             /* OpenBLAS dsyr skips zero entries of x */
            const double t = -xc;
            for (i64 r = 1; r <= c; ++r) PB(j + r, j + c) = fma(t, PB(j, j + r), PB(j + r, j + c));
        }
    }
#undef PB
    return 0;
}

/* DPBTRS: for every right-hand side, DTBSV with U^T then U ('U') or L then L^T ('L') (src/lapack.jl:300-332). */
int oracle_dpbtrs(char uplo, i64 n, i64 kd, i64 nrhs, const double *ab, i64 ldab, double *b, i64 ldb)
{
    const int up = (uplo == 'U' || uplo == 'u');
    for (i64 q = 0; q < nrhs; ++q) {
        double *x = b + q * ldb;
        if (up) {
            oracle_dtbsv('U', 'T', 'N', n, kd, ab, ldab, x);
            oracle_dtbsv('U', 'N', 'N', n, kd, ab, ldab, x);
        } else {
            oracle_dtbsv('L', 'N', 'N', n, kd, ab, ldab, x);
            oracle_dtbsv('L', 'T', 'N', n, kd, ab, ldab, x);
        }
    }
    return 0;
}
