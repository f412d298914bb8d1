/*
 * bmb200_internal.h -- test and tuning hooks exported by libbmb200.so that are NOT part of the
 * drop-in ABI (include/bmb200.h).  Nothing in the reference binds these; they exist so that the
 * parity tests and the A/B timing scripts can reach one kernel variant directly.
 */
#ifndef BMB200_INTERNAL_H
#define BMB200_INTERNAL_H

#include "bmb200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* counts the pairs (x[i], d[i]) for which the fast correctly-rounded division of common.cuh
 * (gb_div) differs from the IEEE quotient x[i] / d[i]; *dbad is a DEVICE counter.            */
int bmb200_internal_divcheck(bmb200_handle_t h, int64_t n, const double *dx, const double *dd,
                             unsigned long long *dbad);

/* same for the verified two-term reciprocal division of the slot-scheduled solve
 * (gbtrs_slot.cu): dbad[0] = pairs whose accepted quotient differs from x/d (must stay 0),
 * dbad[1] = pairs that took the verified fast path (informational).                          */
int bmb200_internal_divcheck2(bmb200_handle_t h, int64_t n, const double *dx, const double *dd,
                              unsigned long long *dbad);

/* bmb200_dgbtrs('N', ...) forced through one variant of the slot-scheduled kernels (gbtrs_slot.cu):
 * PF / PB steps per forward / backward round, W warps per CTA (1, 2, 4), RF / RB right-hand sides
 * per warp in the forward / backward sweep.  -2: variant not instantiated.                      */
int bmb200_internal_gbtrs_slot(bmb200_handle_t h, int PF, int PB, int W, int RF, int RB, int64_t n,
                               int64_t kl, int64_t ku, int64_t nrhs, const double *dAB, int64_t ldab,
                               const int64_t *d_ipiv, double *dB, int64_t ldb);

/* which kernel took the product columns of the last bmb200_dgbmm_bb on this handle: 0 scalar sweep (narrow bands),
 * 1 two-CTA tile kernel (DMMA), 2 persistent ring kernel (DMMA), 3 K-blocked wide-band kernel (DMMA).  smoke() and the
 * tests use it to assert that a wide-band product really ran on the tensor cores. */
int bmb200_internal_last_gbmm_path(bmb200_handle_t h);

/* the generic (typed.cu) LU / solve kernels instantiated for Float64: a cross-check of the tuned Float64 path, never a
 * dispatch target of bmb200_dgbtrf / bmb200_dgbtrs. */
int bmb200_internal_dgbtrf_generic(bmb200_handle_t h, int64_t m, int64_t n, int64_t kl, int64_t ku, double *dAB, int64_t ldab,
                                   int64_t *d_ipiv, int *info);
int bmb200_internal_dgbtrs_generic(bmb200_handle_t h, char trans, int64_t n, int64_t kl, int64_t ku, int64_t nrhs,
                                   const double *dAB, int64_t ldab, const int64_t *d_ipiv, double *dB, int64_t ldb);

/* development knobs of this handle (A/B timing and diagnostics; csrc/common.cuh `bmb_tuning` lists the keys and
 * the shipped defaults; key "reset" restores them).  The library never reads the environment. */
int bmb200_internal_set_tuning(bmb200_handle_t h, const char *key, long long value);

#ifdef __cplusplus
}
#endif
#endif /* BMB200_INTERNAL_H */
