/*
 * bmb200.h -- C ABI of libbmb200.so: the B200 (sm_100a) banded hot path behind
 * BandedMatrices.jl's mul! / * / lu / ldiv! / \ .
 *
 * Every entry point replaces one native call (or one pure-Julia driver loop) of the
 * reference; the citation beside each says which (paths relative to the reference root).
 * The library is what a `ccall((:bmb200_xxx, libbmb200), ...)` in the Julia glue binds
 * (julia/BandedMatricesB200.jl) and what the Python ctypes host mirror
 * (bandedmatrices.jl_b200/) binds in this repository.
 *
 * Conventions
 *   - plain C, by-value scalars, raw pointers; no torch / C++ types cross the boundary.
 *   - all matrices are Float64, column-major, in LAPACK general-band storage:
 *       band storage  A[k,j] at a[(ku + k - j) + j*lda]     (0-based; src/banded/BandedMatrix.jl:414-419)
 *       LU storage    ldab >= 2*kl+ku+1, A[k,j] at ab[(kl + ku + k - j) + j*ldab]
 *   - pointers named d* are DEVICE pointers (HBM resident); h* are HOST pointers.
 *   - ipiv is 1-based int64 exactly as LAPACK/Julia (`Vector{BlasInt}`, src/banded/BandedLU.jl:12).
 *   - return value: 0 ok; -i = the i-th argument (1-based, counting the handle) is invalid
 *     (LAPACK xerbla convention); BMB200_ERR_CUDA (-1000 - cudaError) for runtime failures.
 *     For gbtrf the LAPACK `info` (> 0: first exactly-zero pivot, factorisation completed) is
 *     returned through *info.
 *   - calls are asynchronous on the handle's stream unless they return host-visible data.
 *   - there is NO CPU fallback: a call that cannot run on the device fails loudly.
 */
#ifndef BMB200_H
#define BMB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BMB200_ERR_CUDA (-1000)
#define BMB200_VERSION 100

typedef struct bmb200_ctx *bmb200_handle_t;

/* ---- context / memory (plumbing the Julia glue needs for its device array type) ---- */
int bmb200_version(void);
/* stream may be NULL (legacy default stream) or a cudaStream_t owned by the caller. */
int bmb200_create(bmb200_handle_t *h, int device, void *stream);
int bmb200_destroy(bmb200_handle_t h);
int bmb200_set_stream(bmb200_handle_t h, void *stream);
int bmb200_sync(bmb200_handle_t h);
int bmb200_malloc(bmb200_handle_t h, void **dptr, size_t bytes);
int bmb200_free(bmb200_handle_t h, void *dptr);
int bmb200_memcpy_h2d(bmb200_handle_t h, void *dst, const void *hsrc, size_t bytes);
int bmb200_memcpy_d2h(bmb200_handle_t h, void *hdst, const void *dsrc, size_t bytes);
const char *bmb200_last_error(bmb200_handle_t h);
/* number of kernels this handle has launched since creation (bench.py's gpu_launches) */
int64_t bmb200_launch_count(bmb200_handle_t h);

/* ---- y <- alpha*op(A)*x + beta*y ------------------------------------------------------
 * Replaces dgbmv_ : src/blas.jl:16-28 (pointer form) and BLAS.gbmv! reached from
 * src/generic/matmul.jl:21-23.  Same argument list as the Fortran routine.  trans in
 * {'N','T','C'}.  beta == 0 overwrites y (NaN/Inf in y do not propagate); alpha == 0 only
 * scales; out-of-matrix corner slots of dA are never read.  'N' accumulates every y[i] in
 * ascending-column order with one FMA per term, t = alpha*x[j] rounded first -- the order
 * OpenBLAS' dgbmv_n uses -- so results are bit-identical to the reference CPU path.      */
int bmb200_dgbmv(bmb200_handle_t h, char trans, int64_t m, int64_t n, int64_t kl, int64_t ku,
                 double alpha, const double *dA, int64_t lda, const double *dx, int64_t incx,
                 double beta, double *dy, int64_t incy);

/* ---- C <- alpha*A*B + beta*C, all three banded -----------------------------------------
 * Replaces _gbmm! : src/banded/gbmm.jl:296-340 (the per-column dgbmv_ loop in three regimes
 * plus the trailing beta-fill), one launch instead of m BLAS calls.  A is n x nu with
 * (Al,Au), B is nu x m with (Bl,Bu), C is n x m and dC points at the first WRITTEN band row
 * (gbmm.jl:289) with (Cl,Cu) = min((n-1,m-1),(Al+Bl,Au+Bu)); lda/ldb/ldc are the column
 * strides.  All band widths >= 0 (the gbmm! driver has already pruned negative ones).     */
int bmb200_dgbmm_bb(bmb200_handle_t h, int64_t n, int64_t nu, int64_t m, int64_t Al, int64_t Au,
                    int64_t Bl, int64_t Bu, int64_t Cl, int64_t Cu, double alpha, const double *dA,
                    int64_t lda, const double *dB, int64_t ldb, double beta, double *dC, int64_t ldc);

/* ---- C <- alpha*op(A)*B + beta*C, A banded m x n, B/C dense column-major, nrhs columns ----
 * Replaces the per-column mul! loop of src/generic/matmul.jl:243-256 (one dgbmv_ per column
 * of B): A is streamed once for all right-hand sides.                                      */
int bmb200_dgbmm_bd(bmb200_handle_t h, char trans, int64_t m, int64_t n, int64_t kl, int64_t ku,
                    int64_t nrhs, double alpha, const double *dA, int64_t lda, const double *dB,
                    int64_t ldb, double beta, double *dC, int64_t ldc);

/* ---- C <- alpha*A*op(P) + beta*C, A / C dense column-major (M x K, M x N), P banded -------------
 * Replaces the per-row mul! loop of src/generic/matmul.jl:258-271 (dense x banded: one strided dgbmv_ per ROW of C).
 * trans 'N': op(P) = P, K x N with (kl,ku) (rows of C are dgbmv_('T') dot products: equal to 1e-13);
 * trans 'T': op(P) = P^T, P N x K with (kl,ku) (rows of C are dgbmv_('N'): bit-identical, k ascending).          */
int bmb200_dgbmm_db(bmb200_handle_t h, char trans, int64_t M, int64_t K, int64_t N, int64_t kl, int64_t ku,
                    double alpha, const double *dA, int64_t lda, const double *dP, int64_t ldp, double beta,
                    double *dC, int64_t ldc);

/* ---- C <- beta*C on a rows x cols column-major block; beta == 0 zero-fills ----------------
 * Replaces _fill_lmul!/_fill_rmul! : src/generic/utils.jl:29-31 (used by gbmm.jl:287-288,339
 * and matmul.jl:31-33,45,52).  inc is the element stride inside a column (1 for matrices).  */
int bmb200_dfill_lmul(bmb200_handle_t h, double beta, double *dC, int64_t rows, int64_t cols,
                      int64_t ldc, int64_t inc);

/* ---- widening copy of lu(A): (l+u+1) x n band data -> (2l+u+1) x n, top l rows zero ---------
 * Replaces BandedMatrix{T}(A,(l,l+u)) at src/banded/BandedLU.jl:110 (scalar loop at
 * src/banded/BandedMatrix.jl:222-232).                                                     */
int bmb200_dband_widen(bmb200_handle_t h, int64_t n, int64_t l, int64_t u, const double *dA,
                       int64_t lda, double *dAB, int64_t ldab);

/* ---- partial-pivot band LU -------------------------------------------------------------
 * Replaces dgbtrf_ reached through LAPACK.gbtrf!(kl, ku, m, AB) at src/banded/BandedLU.jl:98.
 * dAB is (ldab >= 2kl+ku+1) x n in LU storage and is overwritten by the factors in LAPACK's
 * format (multipliers not row-permuted, BandedLU.jl:7-8).  d_ipiv: min(m,n) int64 on the
 * DEVICE (1-based); the host copy the reference keeps in BandedLU.ipiv is fetched by the
 * caller with bmb200_memcpy_d2h.  Pivot choice = first maximum of |.|, multipliers scaled by
 * the reciprocal, updates are one FMA per term in ascending column order: pivots AND factors
 * are bit-identical to DGBTF2.  *info (host) receives the LAPACK info; the call synchronises. */
int bmb200_dgbtrf(bmb200_handle_t h, int64_t m, int64_t n, int64_t kl, int64_t ku, double *dAB,
                  int64_t ldab, int64_t *d_ipiv, int *info);

/* ---- solve with the factors ------------------------------------------------------------
 * Replaces dgbtrs_ reached through LAPACK.gbtrs!(trans, kl, ku, m, AB, ipiv, B) at
 * src/banded/linalg.jl:28 ('N'), :46 ('T'), :62 ('C').  dB is n x nrhs, overwritten by X.    */
int bmb200_dgbtrs(bmb200_handle_t h, char trans, int64_t n, int64_t kl, int64_t ku, int64_t nrhs,
                  const double *dAB, int64_t ldab, const int64_t *d_ipiv, double *dB, int64_t ldb);

/* ---- triangular band solve / multiply ---------------------------------
 * Replace dtbsv_ / dtbmv_ reached through tbsv!(uplo, trans, diag, m, k, A, x) (src/blas.jl:109-141) and
 * tbmv!(...) (src/blas.jl:71-105), i.e. ldiv! / lmul! of UpperTriangular / LowerTriangular{<:BandedMatrix}
 * (src/tribanded.jl:47-84).  dA is BLAS triangular-band storage ('U': T[i,j] at dA[(k+i-j) + j*lda],
 * 'L': T[i,j] at dA[(i-j) + j*lda]); dx (n doubles, incx = 1) is overwritten.  trans 'N': bit-identical to OpenBLAS
 * dtbsv_/dtbmv_; 'T'/'C' (dot-product forms, k <= 1024 for dtbsv): equal to 1e-13.                                                  */
int bmb200_dtbsv(bmb200_handle_t h, char uplo, char trans, char diag, int64_t n, int64_t k,
                 const double *dA, int64_t lda, double *dx, int64_t incx);
int bmb200_dtbmv(bmb200_handle_t h, char uplo, char trans, char diag, int64_t n, int64_t k,
                 const double *dA, int64_t lda, double *dx, int64_t incx);

/* ---- symmetric band matvec ----------------------------------------------
 * Replaces dsbmv_ reached through sbmv!(uplo, n, k, alpha, A, lda, x, incx, beta, y, incy) (src/blas.jl:36-66), i.e. mul! of
 * Symmetric{<:BandedMatrix} (src/symbanded/symbanded.jl:72-93).  Only the `uplo` triangle is stored, in triangular-band
 * storage as for dtbsv.  y <- alpha*S*x + beta*y (beta == 0 overwrites); incx = incy = 1; x must not alias y.  Equal to
 * OpenBLAS dsbmv_ to 1e-13 (its summation order is unspecified).                                         */
int bmb200_dsbmv(bmb200_handle_t h, char uplo, int64_t n, int64_t k, double alpha, const double *dA,
                 int64_t lda, const double *dx, int64_t incx, double beta, double *dy, int64_t incy);

/* ---- banded Cholesky (the factorisation half) -----------------------------------------------
 * bmb200_dpbtrf replaces dpbtrf_ reached through pbtrf!(uplo, n, kd, AB) (src/lapack.jl:268-292), i.e. banded_chol! behind
 * cholesky(Symmetric(::BandedMatrix)) (src/symbanded/BandedCholesky.jl:2-13).  dAB is LAPACK symmetric band storage
 * ('U': A[i,k] at dAB[(kd+i-k) + k*ldab], i <= k; 'L': A[i,k] at dAB[(i-k) + k*ldab], i >= k), overwritten by the factor
 * (A = U^T U or L L^T).  *info (host): 0, or j > 0 when the leading minor of order j is not positive definite (the reference
 * turns it into PosDefException); the call synchronises.  kd <= 64 (DPBTF2: the algorithm DPBTRF runs there): bit-identical to
 * OpenBLAS; wider bands run a blocked right-looking factorisation (DPBTRF's DSYRK/DGEMM order is unspecified): equal to
 * rounding (tests: 1e-12 relative, ||U^T U - A|| <= 1e-14 ||A|| kd).
 * bmb200_dpbtrs replaces dpbtrs_ reached through pbtrs!(uplo, n, kd, AB, B) (src/lapack.jl:300-332), i.e. ldiv! of the banded
 * Cholesky factorisation (BandedCholesky.jl:72-80): dB (n x nrhs, column stride ldb) is overwritten by A^{-1} B.            */
int bmb200_dpbtrf(bmb200_handle_t h, char uplo, int64_t n, int64_t kd, double *dAB, int64_t ldab, int *info);
int bmb200_dpbtrs(bmb200_handle_t h, char uplo, int64_t n, int64_t kd, int64_t nrhs, const double *dAB, int64_t ldab,
                  double *dB, int64_t ldb);

/* ---- Float32 / ComplexF32 / ComplexF64 instantiations ------------------------------------------
 * src/blas.jl:4-7 generates gbmv! / sbmv! / hbmv! for the four BLAS element types; LAPACK.gbtrf! / gbtrs! (BandedLU.jl:98,
 * linalg.jl:28,46,62 -- 'C' is a true conjugate-transpose solve) take the same four.  Same argument lists as the d-routines
 * above with typed device pointers behind void* (complex = interleaved re,im as in Julia / Fortran); alpha and beta are passed
 * by HOST pointer to one element of the type (the Fortran convention).  trans 'C' conjugates.  These run one generic kernel
 * per operation (typed.cu) and agree with OpenBLAS to rounding (1e-5 / 1e-13 relative for single / double precision; pivots
 * equal wherever a column's maximum is unique); the tuned pipelines are Float64 only.                                     */
int bmb200_sgbmv(bmb200_handle_t h, char trans, int64_t m, int64_t n, int64_t kl, int64_t ku, const void *alpha,
                 const void *dA, int64_t lda, const void *dx, int64_t incx, const void *beta, void *dy, int64_t incy);
int bmb200_sgbtrf(bmb200_handle_t h, int64_t m, int64_t n, int64_t kl, int64_t ku, void *dAB, int64_t ldab,
                  int64_t *d_ipiv, int *info);
int bmb200_sgbtrs(bmb200_handle_t h, char trans, int64_t n, int64_t kl, int64_t ku, int64_t nrhs, const void *dAB,
                  int64_t ldab, const int64_t *d_ipiv, void *dB, int64_t ldb);
int bmb200_cgbmv(bmb200_handle_t h, char trans, int64_t m, int64_t n, int64_t kl, int64_t ku, const void *alpha,
                 const void *dA, int64_t lda, const void *dx, int64_t incx, const void *beta, void *dy, int64_t incy);
int bmb200_cgbtrf(bmb200_handle_t h, int64_t m, int64_t n, int64_t kl, int64_t ku, void *dAB, int64_t ldab,
                  int64_t *d_ipiv, int *info);
int bmb200_cgbtrs(bmb200_handle_t h, char trans, int64_t n, int64_t kl, int64_t ku, int64_t nrhs, const void *dAB,
                  int64_t ldab, const int64_t *d_ipiv, void *dB, int64_t ldb);
int bmb200_zgbmv(bmb200_handle_t h, char trans, int64_t m, int64_t n, int64_t kl, int64_t ku, const void *alpha,
                 const void *dA, int64_t lda, const void *dx, int64_t incx, const void *beta, void *dy, int64_t incy);
int bmb200_zgbtrf(bmb200_handle_t h, int64_t m, int64_t n, int64_t kl, int64_t ku, void *dAB, int64_t ldab,
                  int64_t *d_ipiv, int *info);
int bmb200_zgbtrs(bmb200_handle_t h, char trans, int64_t n, int64_t kl, int64_t ku, int64_t nrhs, const void *dAB,
                  int64_t ldab, const int64_t *d_ipiv, void *dB, int64_t ldb);
/* the remaining S / C / Z entry points of the path (tbsv! / tbmv! src/blas.jl:71-141, pbtrf! / pbtrs! src/lapack.jl:268-332 with
 * Hermitian semantics for the complex types, _gbmm! src/banded/gbmm.jl:296-340 and the banded x dense loop
 * src/generic/matmul.jl:243-256): same argument lists as the d-routines, generic kernels (typed.cu), equal to OpenBLAS to rounding. */
int bmb200_stbsv(bmb200_handle_t h, char uplo, char trans, char diag, int64_t n, int64_t k, const void *dA, int64_t lda,
                 void *dx, int64_t incx);
int bmb200_stbmv(bmb200_handle_t h, char uplo, char trans, char diag, int64_t n, int64_t k, const void *dA, int64_t lda,
                 void *dx, int64_t incx);
int bmb200_spbtrf(bmb200_handle_t h, char uplo, int64_t n, int64_t kd, void *dAB, int64_t ldab, int *info);
int bmb200_spbtrs(bmb200_handle_t h, char uplo, int64_t n, int64_t kd, int64_t nrhs, const void *dAB, int64_t ldab, void *dB,
                  int64_t ldb);
int bmb200_sgbmm_bb(bmb200_handle_t h, int64_t n, int64_t nu, int64_t m, int64_t Al, int64_t Au, int64_t Bl, int64_t Bu,
                    int64_t Cl, int64_t Cu, const void *alpha, const void *dA, int64_t lda, const void *dB, int64_t ldb,
                    const void *beta, void *dC, int64_t ldc);
int bmb200_sgbmm_bd(bmb200_handle_t h, char trans, int64_t m, int64_t n, int64_t kl, int64_t ku, int64_t nrhs,
                    const void *alpha, const void *dA, int64_t lda, const void *dB, int64_t ldb, const void *beta, void *dC,
                    int64_t ldc);
int bmb200_ctbsv(bmb200_handle_t h, char uplo, char trans, char diag, int64_t n, int64_t k, const void *dA, int64_t lda,
                 void *dx, int64_t incx);
int bmb200_ctbmv(bmb200_handle_t h, char uplo, char trans, char diag, int64_t n, int64_t k, const void *dA, int64_t lda,
                 void *dx, int64_t incx);
int bmb200_cpbtrf(bmb200_handle_t h, char uplo, int64_t n, int64_t kd, void *dAB, int64_t ldab, int *info);
int bmb200_cpbtrs(bmb200_handle_t h, char uplo, int64_t n, int64_t kd, int64_t nrhs, const void *dAB, int64_t ldab, void *dB,
                  int64_t ldb);
int bmb200_cgbmm_bb(bmb200_handle_t h, int64_t n, int64_t nu, int64_t m, int64_t Al, int64_t Au, int64_t Bl, int64_t Bu,
                    int64_t Cl, int64_t Cu, const void *alpha, const void *dA, int64_t lda, const void *dB, int64_t ldb,
                    const void *beta, void *dC, int64_t ldc);
int bmb200_cgbmm_bd(bmb200_handle_t h, char trans, int64_t m, int64_t n, int64_t kl, int64_t ku, int64_t nrhs,
                    const void *alpha, const void *dA, int64_t lda, const void *dB, int64_t ldb, const void *beta, void *dC,
                    int64_t ldc);
int bmb200_ztbsv(bmb200_handle_t h, char uplo, char trans, char diag, int64_t n, int64_t k, const void *dA, int64_t lda,
                 void *dx, int64_t incx);
int bmb200_ztbmv(bmb200_handle_t h, char uplo, char trans, char diag, int64_t n, int64_t k, const void *dA, int64_t lda,
                 void *dx, int64_t incx);
int bmb200_zpbtrf(bmb200_handle_t h, char uplo, int64_t n, int64_t kd, void *dAB, int64_t ldab, int *info);
int bmb200_zpbtrs(bmb200_handle_t h, char uplo, int64_t n, int64_t kd, int64_t nrhs, const void *dAB, int64_t ldab, void *dB,
                  int64_t ldb);
int bmb200_zgbmm_bb(bmb200_handle_t h, int64_t n, int64_t nu, int64_t m, int64_t Al, int64_t Au, int64_t Bl, int64_t Bu,
                    int64_t Cl, int64_t Cu, const void *alpha, const void *dA, int64_t lda, const void *dB, int64_t ldb,
                    const void *beta, void *dC, int64_t ldc);
int bmb200_zgbmm_bd(bmb200_handle_t h, char trans, int64_t m, int64_t n, int64_t kl, int64_t ku, int64_t nrhs,
                    const void *alpha, const void *dA, int64_t lda, const void *dB, int64_t ldb, const void *beta, void *dC,
                    int64_t ldc);
/* symmetric (Float32) / Hermitian (complex) band matvec from one stored triangle: ssbmv_ / chbmv_ / zhbmv_ reached through
 * sbmv! / hbmv! (src/blas.jl:36-66), i.e. mul! of Symmetric / Hermitian{<:BandedMatrix} (src/symbanded/symbanded.jl:72-96).
 * Only the real part of the diagonal is read, as in xHBMV.  incx = incy = 1; x must not alias y.                            */
int bmb200_ssbmv(bmb200_handle_t h, char uplo, int64_t n, int64_t k, const void *alpha, const void *dA, int64_t lda,
                 const void *dx, int64_t incx, const void *beta, void *dy, int64_t incy);
int bmb200_chbmv(bmb200_handle_t h, char uplo, int64_t n, int64_t k, const void *alpha, const void *dA, int64_t lda,
                 const void *dx, int64_t incx, const void *beta, void *dy, int64_t incy);
int bmb200_zhbmv(bmb200_handle_t h, char uplo, int64_t n, int64_t k, const void *alpha, const void *dA, int64_t lda,
                 const void *dx, int64_t incx, const void *beta, void *dy, int64_t incy);

/* ---- band-aligned elementwise operations between different bandwidths ----
 * bmb200_dband_axpy replaces banded_axpy!(a, X, Y) (src/banded/BandedMatrix.jl:1006-1015 -> axpy!(a, X.data, Y.data) for equal
 * bandwidths: one FMA per slot, as OpenBLAS daxpy; src/generic/broadcast.jl:978-1020 otherwise: Y[k,j] = a*X[k,j] + Y[k,j] on
 * the overlapping bands).  bmb200_dband_copy replaces copyto!(dest, src) between bandwidths (broadcast.jl:175-230): overlapping
 * bands copied, dest's other in-matrix band entries zeroed.  *nonzero_outside (host) = number of non-zero entries of X / src in
 * bands the destination does not store; if it is not 0 nothing was written and the caller raises BandError.  The calls that
 * have to look (unequal bandwidths) synchronise.                                                          */
int bmb200_dband_axpy(bmb200_handle_t h, int64_t m, int64_t n, double a, int64_t xl, int64_t xu, const double *dX,
                      int64_t ldx, int64_t yl, int64_t yu, double *dY, int64_t ldy, int64_t *nonzero_outside);
int bmb200_dband_copy(bmb200_handle_t h, int64_t m, int64_t n, int64_t sl, int64_t su, const double *dS, int64_t lds,
                      int64_t dl, int64_t du, double *dD, int64_t ldd, int64_t *nonzero_outside);

/* ---- band utilities of the gbmm! driver and the broadcasting layer (all on the device) --------------------------------
 * bmb200_dband_lmul_block : lmul!(beta, view(C, r0+1:r1, c0+1:c1)) on the stored band, beta == 0 zero-fills -- the blocks a
 *                           negative-bandwidth operand leaves untouched in gbmm! (src/banded/gbmm.jl:234-249).
 * bmb200_dband_transpose  : D = A' as a plain BandedMatrix (convert at src/generic/matmul.jl:182-184): A is m x n with (l,u),
 *                           D is n x m with (u,l); corner slots of D are zeroed.
 * bmb200_dband_nonzero_rows: flags_host[r] = 1 when band row r of the data array holds a non-zero in-matrix entry
 *                           (gbmm.jl:191-205 counts the leading / trailing all-zero bands from these); synchronises.
 * bmb200_dband_axpby      : Z = alpha*X + beta*Y entry by entry over Z's band, X / Y read as zero outside their bands --
 *                           the arithmetic of the reference's broadcast kernels for A .+ B, A .- B, a .* A, a .* A .+ b .* B
 *                           (src/generic/broadcast.jl:359-384, 927-964; products and sum rounded separately).  Z may alias X or Y
 *                           when the bandwidths are equal.                                                                        */
int bmb200_dband_lmul_block(bmb200_handle_t h, int64_t m, int64_t n, int64_t l, int64_t u, double *dC, int64_t ldc,
                            int64_t r0, int64_t r1, int64_t c0, int64_t c1, double beta);
int bmb200_dband_transpose(bmb200_handle_t h, int64_t m, int64_t n, int64_t l, int64_t u, const double *dS, int64_t lds,
                           double *dD, int64_t ldd);
int bmb200_dband_nonzero_rows(bmb200_handle_t h, int64_t m, int64_t n, int64_t l, int64_t u, const double *dX, int64_t ldx,
                              int *flags_host);
int bmb200_dband_axpby(bmb200_handle_t h, int64_t m, int64_t n, double alpha, int64_t xl, int64_t xu, const double *dX,
                       int64_t ldx, double beta, int64_t yl, int64_t yu, const double *dY, int64_t ldy, int64_t zl,
                       int64_t zu, double *dZ, int64_t ldz);

/* ---- lu(A): widening copy + factorisation in one call ------------------------------------------
 * Replaces _lu, src/banded/BandedLU.jl:106-111 ( lu!(BandedMatrix{T}(A,(l,l+u))) ): dA is the (kl+ku+1) x n band storage of
 * A, dAB receives the (2kl+ku+1) x n LU storage.  Same result as bmb200_dband_widen followed by bmb200_dgbtrf; giving the
 * source lets the interchange-free wide-band path skip its own copy of the band.                                      */
int bmb200_dgbtrf_from(bmb200_handle_t h, int64_t m, int64_t n, int64_t kl, int64_t ku, const double *dA, int64_t lda,
                       double *dAB, int64_t ldab, int64_t *d_ipiv, int *info);

/* ---- host-buffer forms: what a Fortran-ABI caller with HOST arrays gets (bench.py "e2e") ----
 * Same semantics as the calls above; inputs are copied host->device in pipelined chunks, the
 * result is copied back, and the call returns after the result is in host memory.           */
int bmb200_dgbmv_host(bmb200_handle_t h, char trans, int64_t m, int64_t n, int64_t kl, int64_t ku,
                      double alpha, const double *hA, int64_t lda, const double *hx, int64_t incx,
                      double beta, double *hy, int64_t incy);
int bmb200_dgbsv_host(bmb200_handle_t h, int64_t n, int64_t kl, int64_t ku, int64_t nrhs,
                      double *hAB, int64_t ldab, int64_t *h_ipiv, double *hB, int64_t ldb, int *info);
int bmb200_dgbmm_bb_host(bmb200_handle_t h, int64_t n, int64_t nu, int64_t m, int64_t Al, int64_t Au,
                         int64_t Bl, int64_t Bu, int64_t Cl, int64_t Cu, double alpha, const double *hA,
                         int64_t lda, const double *hB, int64_t ldb, double beta, double *hC, int64_t ldc);

/* ---- multi-GPU: row-sharded gbmv with an (l+u) halo of x over NVLink peer memory ------------
 * One process per GPU.  Rank r owns rows/columns [c0, c1) of a square
 * n x n matrix: dA_local holds data columns [c0, c1), dx_local / dy_local the matching slices.
 * bmb200_halo_* wires the per-rank halo mailboxes (CUDA IPC handles exchanged by the host
 * plumbing, e.g. torch.distributed); bmb200_dgbmv_sharded pushes the kl/ku boundary entries of
 * x into the neighbours' mailboxes with peer stores, signals, waits for its own halo and runs
 * the same streaming kernel on the local slab.  No reduction, no NCCL on the data path.       */
#define BMB200_IPC_HANDLE_BYTES 64
int bmb200_halo_create(bmb200_handle_t h, int64_t max_halo, void *ipc_handle_out /* 64 bytes */);
int bmb200_halo_connect(bmb200_handle_t h, int rank, int nranks, const void *ipc_left /* or NULL */,
                        const void *ipc_right /* or NULL */);
int bmb200_halo_destroy(bmb200_handle_t h);
int bmb200_dgbmv_sharded(bmb200_handle_t h, int64_t n_global, int64_t c0, int64_t c1, int64_t kl,
                         int64_t ku, double alpha, const double *dA_local, int64_t lda,
                         const double *dx_local, double beta, double *dy_local);

#ifdef __cplusplus
}
#endif
#endif /* BMB200_H */
