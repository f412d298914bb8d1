/*
 * blasdriver.c -- CPU-BASELINE driver.  TEST / BENCH INFRASTRUCTURE ONLY (see bmoracle.c header).
 *
 * dlopen()s the OpenBLAS ILP64 library bundled with numpy and replays, from C, the exact
 * call sequences BandedMatrices.jl issues on the CPU, so that the reference CPU path can be
 * timed beside the GPU numbers without Python call overhead:
 *   drv_gbmv   : one dgbmv_          (src/generic/matmul.jl:21-23)
 *   drv_gbmm   : _gbmm! -- one dgbmv_ per column of C in three regimes + trailing beta fill
 *                (src/banded/gbmm.jl:296-340; 2^22 tiny calls at config C3)
 *   drv_gbtrf  : dgbtrf_             (src/banded/BandedLU.jl:98)
 *   drv_gbtrs  : dgbtrs_             (src/banded/linalg.jl:28)
 * Every function returns the elapsed wall time in seconds (CLOCK_MONOTONIC) of the BLAS work.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <time.h>

typedef int64_t i64;
typedef void (*gbmv_t)(const char *, const i64 *, const i64 *, const i64 *, const i64 *, const double *,
                       const double *, const i64 *, const double *, const i64 *, const double *, double *,
                       const i64 *, long);
typedef void (*gbtrf_t)(const i64 *, const i64 *, const i64 *, const i64 *, double *, const i64 *, i64 *, i64 *);
typedef void (*gbtrs_t)(const char *, const i64 *, const i64 *, const i64 *, const i64 *, const double *,
                        const i64 *, const i64 *, double *, const i64 *, i64 *, long);
typedef void (*tb_t)(const char *, const char *, const char *, const i64 *, const i64 *, const double *, const i64 *, double *,
                     const i64 *, long, long, long);
typedef void (*sbmv_t)(const char *, const i64 *, const i64 *, const double *, const double *, const i64 *, const double *, const i64 *,
                       const double *, double *, const i64 *, long);
typedef void (*setthr_t)(int);

static void *lib;
static gbmv_t f_gbmv;
static gbtrf_t f_gbtrf;
static gbtrs_t f_gbtrs;
static setthr_t f_setthr;
static tb_t f_tbsv, f_tbmv;
static sbmv_t f_sbmv;

static double now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
static inline i64 imin(i64 a, i64 b) { return a < b ? a : b; }

int drv_open(const char *path)
{
    lib = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!lib) { fprintf(stderr, "blasdriver: %s\n", dlerror()); return -1; }
    f_gbmv = (gbmv_t)dlsym(lib, "scipy_dgbmv_64_");
    f_gbtrf = (gbtrf_t)dlsym(lib, "scipy_dgbtrf_64_");
    f_gbtrs = (gbtrs_t)dlsym(lib, "scipy_dgbtrs_64_");
    f_setthr = (setthr_t)dlsym(lib, "scipy_openblas_set_num_threads64_");
    f_tbsv = (tb_t)dlsym(lib, "scipy_dtbsv_64_");
    f_tbmv = (tb_t)dlsym(lib, "scipy_dtbmv_64_");
    f_sbmv = (sbmv_t)dlsym(lib, "scipy_dsbmv_64_");
    return (f_gbmv && f_gbtrf && f_gbtrs && f_setthr && f_tbsv && f_tbmv && f_sbmv) ? 0 : -2;
}

void drv_set_threads(int k) { f_setthr(k); }

static inline void gbmv(i64 m, i64 n, i64 kl, i64 ku, double alpha, const double *a, i64 lda, const double *x,
                        double beta, double *y)
{
    const i64 one = 1;
    f_gbmv("N", &m, &n, &kl, &ku, &alpha, a, &lda, x, &one, &beta, y, &one, 1);
}

double drv_gbmv(i64 m, i64 n, i64 kl, i64 ku, double alpha, const double *a, i64 lda, const double *x,
                double beta, double *y)
{
    double t0 = now();
    gbmv(m, n, kl, ku, alpha, a, lda, x, beta, y);
    return now() - t0;
}

double drv_gbmm(i64 n, i64 nu, i64 m, i64 Al, i64 Au, i64 Bl, i64 Bu, i64 Cl, i64 Cu, double alpha,
                const double *a, i64 sta, const double *b, i64 stb, double beta, double *c, i64 stc)
{
    double t0 = now();
    i64 j;
    for (j = 1; j <= imin(m, 1 + Bu); ++j) /* gbmm.jl:306-313 */
        gbmv(imin(Cl + j, n), imin(Bl + j, nu), Al, Au, alpha, a, sta, b + ((j - 1) * stb + Bu - j + 1), beta,
             c + ((j - 1) * stc + Cu - j + 1));
    for (j = 2 + Bu; j <= imin(imin(1 + Cu, nu + Bu), m); ++j) /* gbmm.jl:318-325 */
        gbmv(imin(Cl + j, n), imin(Bl + Bu + 1, nu - j + Bu + 1), Al + j - Bu - 1, Au - j + Bu + 1, alpha,
             a + (j - Bu - 1) * sta, sta, b + (j - 1) * stb, beta, c + ((j - 1) * stc + Cu - j + 1));
    for (j = 2 + Cu; j <= imin(imin(m, nu + Bu), n + Cu); ++j) /* gbmm.jl:329-336 */
        gbmv(imin(Cl + Cu + 1, n - j + Cu + 1), imin(Bl + Bu + 1, nu - (j - Bu) + 1), Al + Au, 0, alpha,
             a + (j - Bu - 1) * sta, sta, b + (j - 1) * stb, beta, c + (j - 1) * stc);
    for (j = nu + Bu + 1; j <= imin(m, n + Cu); ++j) /* gbmm.jl:339 */
        for (i64 r = 0; r < Cl + Cu + 1; ++r)
            c[r + (j - 1) * stc] = (beta == 0.0) ? 0.0 : beta * c[r + (j - 1) * stc];
    return now() - t0;
}

double drv_gbtrf(i64 m, i64 n, i64 kl, i64 ku, double *ab, i64 ldab, i64 *ipiv, i64 *info)
{
    double t0 = now();
    f_gbtrf(&m, &n, &kl, &ku, ab, &ldab, ipiv, info);
    return now() - t0;
}

double drv_gbtrs(i64 n, i64 kl, i64 ku, i64 nrhs, const double *ab, i64 ldab, const i64 *ipiv, double *b, i64 ldb,
                 i64 *info)
{
    double t0 = now();
    f_gbtrs("N", &n, &kl, &ku, &nrhs, ab, &ldab, ipiv, b, &ldb, info, 1);
    return now() - t0;
}

/* dtbsv_ (which = 0) / dtbmv_ (which = 1), trans 'N', as tbsv! / tbmv! call them (src/blas.jl:94-99, :132-137) */
double drv_tb(int which, char uplo, char diag, i64 n, i64 k, const double *a, i64 lda, double *x)
{
    const i64 one = 1;
    double t0 = now();
    (which ? f_tbmv : f_tbsv)(&uplo, "N", &diag, &n, &k, a, &lda, x, &one, 1, 1, 1);
    return now() - t0;
}

/* dsbmv_ as sbmv! calls it (src/blas.jl:51-60) */
double drv_sbmv(char uplo, i64 n, i64 k, double alpha, const double *a, i64 lda, const double *x, double beta, double *y)
{
    const i64 one = 1;
    double t0 = now();
    f_sbmv(&uplo, &n, &k, &alpha, a, &lda, x, &one, &beta, y, &one, 1);
    return now() - t0;
}
