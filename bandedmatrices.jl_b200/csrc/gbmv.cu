// gbmv.cu -- y <- alpha*op(A)*x + beta*y for LAPACK band storage on sm_100a.
//
// Replaces dgbmv_ (src/blas.jl:16-28; BLAS.gbmv! from src/generic/matmul.jl:21-23).
//
// Kernels
//   gbmv_n_systolic<W,LDV>  narrow bands (W = kl+ku+1 <= 16), unit strides.  HBM-streaming design:
//       lane c of a warp owns band-storage column c (64 contiguous bytes when lda == 8, fetched with
//       two 256-bit LDG.E.256), x[c] and y are touched exactly once with coalesced accesses, and the
//       W partial sums of every output row travel lane -> lane+1 with one warp shuffle per diagonal
//       (a systolic chain), so each y[i] is accumulated in ASCENDING COLUMN ORDER with one FMA per
//       term -- bit-identical to OpenBLAS' dgbmv_n -- and no shared memory, no re-reads and no
//       bank conflicts are involved.  Algorithmic bytes per row: 8*(lda + 2) (+8 when beta != 0).
//   gbmv_n_sweep<RPL>       any kl, ku, lda, incx, incy: lane = row, sweep over columns; the
//       loads of one column are contiguous over lanes.  Efficient when W >~ 32.
//   gbmv_t_lane<W,LDV>      'T', narrow: lane = column, dot with the x window.
//   gbmv_t_warp             'T', any width: warp = column, shuffle reduction.
#include "common.cuh"
#include "gbmv_systolic.cuh"

// ------------------------------------------------------------------------------------------------
// Narrow-band streaming kernel ('N'): body in gbmv_systolic.cuh
// ------------------------------------------------------------------------------------------------
template <int W, int LDV>
__global__ void __launch_bounds__(256)
gbmv_n_systolic(i64 m, i64 n, int kl, int ku, double alpha, const double *__restrict__ a, i64 lda,
                const double *__restrict__ x, double beta, double *__restrict__ y, i64 total_sets,
                i64 sets_per_run, i64 num_runs)
{
    gbmv_n_systolic_body<W, LDV, XPlain>(m, n, kl, ku, alpha, a, lda, XPlain{x}, beta, y, total_sets, sets_per_run,
                                         num_runs);
}

// ------------------------------------------------------------------------------------------------
// Generic 'N' kernel: lane = row, sweep over columns (ascending => same FMA order as dgbmv_n).
// ------------------------------------------------------------------------------------------------
template <int RPL>
__global__ void __launch_bounds__(256)
gbmv_n_sweep(i64 m, i64 n, i64 kl, i64 ku, double alpha, const double *__restrict__ a, i64 lda,
             const double *__restrict__ x, i64 incx, double beta, double *__restrict__ y, i64 incy, i64 ntiles)
{
    const int lane = threadIdx.x & 31;
    const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const i64 nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    for (i64 tile = warp; tile < ntiles; tile += nwarps) {
        const i64 i0 = tile * (32 * RPL);
        double acc[RPL];
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            const i64 i = i0 + r * 32 + lane;
            acc[r] = (beta == 0.0 || i >= m) ? 0.0 : __dmul_rn(beta, y[i * incy]);
        }
        i64 jlo = i0 - kl;
        if (jlo < 0) jlo = 0;
        i64 jhi = i0 + 32 * RPL - 1 + ku;
        if (jhi > n - 1) jhi = n - 1;
#pragma unroll 4
        for (i64 j = jlo; j <= jhi; ++j) {
            const double t = __dmul_rn(alpha, x[j * incx]);
            const double *colp = a + j * lda + (ku - j);
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const i64 i = i0 + r * 32 + lane;
                if (i < m && i >= j - ku && i <= j + kl) acc[r] = fma(t, colp[i], acc[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            const i64 i = i0 + r * 32 + lane;
            if (i < m) y[i * incy] = acc[r];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// 'T': y[j] = beta*y[j] + alpha * sum_i A[i,j]*x[i]
// ------------------------------------------------------------------------------------------------
template <int W, int LDV>
__global__ void __launch_bounds__(256)
gbmv_t_lane(i64 m, i64 n, int kl, int ku, double alpha, const double *__restrict__ a, i64 lda,
            const double *__restrict__ x, double beta, double *__restrict__ y, i64 total_sets)
{
    const int lane = threadIdx.x & 31;
    const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const i64 nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    for (i64 set = warp; set < total_sets; set += nwarps) {
        const i64 c = set * 32 + lane;
        const bool valid = c < n;
        double col[W];
        load_col<W, LDV>(a, lda, c, valid, col);
        double temp = 0.0;
#pragma unroll
        for (int r = 0; r < W; ++r) {
            const i64 i = c - ku + r;
            if (valid && i >= 0 && i < m) temp = fma(col[r], x[i], temp);
        }
        if (valid) {
            const double y0 = (beta == 0.0) ? 0.0 : __dmul_rn(beta, y[c]);
            y[c] = fma(alpha, temp, y0);
        }
    }
}

__global__ void __launch_bounds__(256)
gbmv_t_warp(i64 m, i64 n, i64 kl, i64 ku, double alpha, const double *__restrict__ a, i64 lda,
            const double *__restrict__ x, i64 incx, double beta, double *__restrict__ y, i64 incy)
{
    const int lane = threadIdx.x & 31;
    const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const i64 nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    for (i64 j = warp; j < n; j += nwarps) {
        i64 i0 = j - ku;
        if (i0 < 0) i0 = 0;
        i64 i1 = j + kl;
        if (i1 > m - 1) i1 = m - 1;
        const double *colp = a + j * lda + (ku - j);
        double temp = 0.0;
        for (i64 i = i0 + lane; i <= i1; i += 32) temp = fma(colp[i], x[i * incx], temp);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) temp += __shfl_xor_sync(0xffffffffu, temp, o);
        if (lane == 0) {
            const double y0 = (beta == 0.0) ? 0.0 : __dmul_rn(beta, y[j * incy]);
            y[j * incy] = fma(alpha, temp, y0);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host dispatch
// ------------------------------------------------------------------------------------------------
template <int W, int LDV>
static int launch_systolic(bmb200_ctx *h, i64 m, i64 n, i64 kl, i64 ku, double alpha, const double *dA, i64 lda,
                           const double *dx, double beta, double *dy)
{
    const int threads = 256;
    int per_sm = 0;
    BMB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gbmv_n_systolic<W, LDV>, threads, 0));
    const SystolicPlan p = systolic_plan(m, ku, h->sm_count, per_sm, threads, h->tune.gbmv_spr);
    gbmv_n_systolic<W, LDV><<<(unsigned)p.blocks, threads, 0, h->stream>>>(m, n, (int)kl, (int)ku, alpha, dA, lda, dx,
                                                                           beta, dy, p.total_sets, p.sets_per_run,
                                                                           p.num_runs);
    BMB_LAUNCH_CHECK(h);
    return 0;
}

template <int W, int LDV>
static int launch_t_lane(bmb200_ctx *h, i64 m, i64 n, i64 kl, i64 ku, double alpha, const double *dA, i64 lda,
                         const double *dx, double beta, double *dy)
{
    const i64 total_sets = cdiv64(n, 32);
    const int threads = 256;
    i64 blocks = imin64(cdiv64(total_sets, threads / 32), (i64)h->sm_count * 8);
    gbmv_t_lane<W, LDV><<<(unsigned)blocks, threads, 0, h->stream>>>(m, n, (int)kl, (int)ku, alpha, dA, lda, dx, beta,
                                                                     dy, total_sets);
    BMB_LAUNCH_CHECK(h);
    return 0;
}

#define BMB_W_CASES(FN, LDV)                        \
    switch (W) {                                    \
    case 1: return FN<1, LDV>(h, m, n, kl, ku, alpha, dA, lda, dx, beta, dy);   \
    case 2: return FN<2, LDV>(h, m, n, kl, ku, alpha, dA, lda, dx, beta, dy);   \
    case 3: return FN<3, LDV>(h, m, n, kl, ku, alpha, dA, lda, dx, beta, dy);   \
    case 4: return FN<4, LDV>(h, m, n, kl, ku, alpha, dA, lda, dx, beta, dy);   \
    case 5: return FN<5, LDV>(h, m, n, kl, ku, alpha, dA, lda, dx, beta, dy);   \
    case 6: return FN<6, LDV>(h, m, n, kl, ku, alpha, dA, lda, dx, beta, dy);   \
    case 7: return FN<7, LDV>(h, m, n, kl, ku, alpha, dA, lda, dx, beta, dy);   \
    case 8: return FN<8, LDV>(h, m, n, kl, ku, alpha, dA, lda, dx, beta, dy);   \
    default: break;                                 \
    }

#define BMB_W_CASES_HI(FN, LDV)                     \
    switch (W) {                                    \
    case 9: return FN<9, LDV>(h, m, n, kl, ku, alpha, dA, lda, dx, beta, dy);   \
    case 10: return FN<10, LDV>(h, m, n, kl, ku, alpha, dA, lda, dx, beta, dy); \
    case 11: return FN<11, LDV>(h, m, n, kl, ku, alpha, dA, lda, dx, beta, dy); \
    case 12: return FN<12, LDV>(h, m, n, kl, ku, alpha, dA, lda, dx, beta, dy); \
    case 13: return FN<13, LDV>(h, m, n, kl, ku, alpha, dA, lda, dx, beta, dy); \
    case 14: return FN<14, LDV>(h, m, n, kl, ku, alpha, dA, lda, dx, beta, dy); \
    case 15: return FN<15, LDV>(h, m, n, kl, ku, alpha, dA, lda, dx, beta, dy); \
    case 16: return FN<16, LDV>(h, m, n, kl, ku, alpha, dA, lda, dx, beta, dy); \
    default: break;                                 \
    }

static int dispatch_narrow_n(bmb200_ctx *h, i64 m, i64 n, i64 kl, i64 ku, double alpha, const double *dA, i64 lda,
                             const double *dx, double beta, double *dy)
{
    const int W = (int)(kl + ku + 1);
    const uintptr_t ap = (uintptr_t)dA;
    if (lda == 8 && (ap & 31) == 0) { BMB_W_CASES(launch_systolic, 8) }
    if (lda == 4 && (ap & 31) == 0 && W <= 4) { BMB_W_CASES(launch_systolic, 4) }
    if (lda == 2 && (ap & 15) == 0 && W <= 2) { BMB_W_CASES(launch_systolic, 2) }
    BMB_W_CASES(launch_systolic, 0)
    BMB_W_CASES_HI(launch_systolic, 0)
    return 1;  // not handled
}

static int dispatch_narrow_t(bmb200_ctx *h, i64 m, i64 n, i64 kl, i64 ku, double alpha, const double *dA, i64 lda,
                             const double *dx, double beta, double *dy)
{
    const int W = (int)(kl + ku + 1);
    const uintptr_t ap = (uintptr_t)dA;
    if (lda == 8 && (ap & 31) == 0) { BMB_W_CASES(launch_t_lane, 8) }
    if (lda == 4 && (ap & 31) == 0 && W <= 4) { BMB_W_CASES(launch_t_lane, 4) }
    BMB_W_CASES(launch_t_lane, 0)
    BMB_W_CASES_HI(launch_t_lane, 0)
    return 1;
}

// Validates like DGBMV's xerbla block (argument positions count the handle as #1).
static int gbmv_check(bmb200_handle_t h, char trans, i64 m, i64 n, i64 kl, i64 ku, const double *dA, i64 lda,
                      const double *dx, i64 incx, double *dy, i64 incy, bool *tr)
{
    if (!h) return -1;
    *tr = (trans == 'T' || trans == 't' || trans == 'C' || trans == 'c');
    if (!*tr && !(trans == 'N' || trans == 'n')) return -2;
    if (m < 0) return -3;
    if (n < 0) return -4;
    if (kl < 0) return -5;
    if (ku < 0) return -6;
    if (lda < kl + ku + 1) return -9;
    if (incx == 0) return -11;
    if (incy == 0) return -14;
    if (m > 0 && n > 0 && (!dA || !dx || !dy)) return -8;
    return 0;
}

// x0 / y0 point at ELEMENT 0 of x / y (incx, incy may be negative).
int bmb_gbmv_device(bmb200_ctx *h, bool tr, i64 m, i64 n, i64 kl, i64 ku, double alpha, const double *dA, i64 lda,
                    const double *x0, i64 incx, double beta, double *y0, i64 incy)
{
    if (m == 0 || n == 0) return 0;  // DGBMV quick return
    const i64 leny = tr ? n : m;
    if (alpha == 0.0) {  // y <- beta*y only; A and x are not referenced
        if (beta == 1.0) return 0;
        double *ylow = incy > 0 ? y0 : y0 + (leny - 1) * incy;
        return bmb200_dfill_lmul(h, beta, ylow, leny, 1, 0, incy < 0 ? -incy : incy);
    }
    // clamp band widths to what the matrix can hold (keeps template widths small for tiny matrices)
    const i64 kle = imin64(kl, m - 1), kue = imin64(ku, n - 1);
    const double *ae = dA + (ku - kue);  // row offset so that A[k,j] stays at ae[(kue+k-j) + j*lda]
    const bool unit = (incx == 1 && incy == 1);
    if (unit && kle + kue + 1 <= 16) {
        int r = tr ? dispatch_narrow_t(h, m, n, kle, kue, alpha, ae, lda, x0, beta, y0)
                   : dispatch_narrow_n(h, m, n, kle, kue, alpha, ae, lda, x0, beta, y0);
        if (r <= 0) return r;
    }
    const int threads = 256;
    if (!tr) {
        const bool wide = (kle + kue + 1) >= 48;
        const i64 rows_per_tile = wide ? 64 : 32;
        const i64 ntiles = cdiv64(m, rows_per_tile);
        const i64 blocks = imin64(cdiv64(ntiles, threads / 32), (i64)h->sm_count * 8);
        if (wide)
            gbmv_n_sweep<2><<<(unsigned)blocks, threads, 0, h->stream>>>(m, n, kle, kue, alpha, ae, lda, x0, incx, beta,
                                                                         y0, incy, ntiles);
        else
            gbmv_n_sweep<1><<<(unsigned)blocks, threads, 0, h->stream>>>(m, n, kle, kue, alpha, ae, lda, x0, incx, beta,
                                                                         y0, incy, ntiles);
    } else {
        const i64 blocks = imin64(cdiv64(n, threads / 32), (i64)h->sm_count * 8);
        gbmv_t_warp<<<(unsigned)blocks, threads, 0, h->stream>>>(m, n, kle, kue, alpha, ae, lda, x0, incx, beta, y0,
                                                                 incy);
    }
    BMB_LAUNCH_CHECK(h);
    return 0;
}

extern "C" int bmb200_dgbmv(bmb200_handle_t h, char trans, int64_t m, int64_t n, int64_t kl, int64_t ku, double alpha,
                            const double *dA, int64_t lda, const double *dx, int64_t incx, double beta, double *dy,
                            int64_t incy)
{
    bool tr;
    int rc = gbmv_check(h, trans, m, n, kl, ku, dA, lda, dx, incx, dy, incy, &tr);
    if (rc) return rc;
    DeviceGuard g(h->device);
    // BLAS convention: with a negative increment the vector is stored backwards from the pointer given
    const i64 lenx = tr ? m : n, leny = tr ? n : m;
    const double *x0 = incx > 0 ? dx : dx - (lenx - 1) * incx;
    double *y0 = incy > 0 ? dy : dy - (leny - 1) * incy;
    return bmb_gbmv_device(h, tr, m, n, kl, ku, alpha, dA, lda, x0, incx, beta, y0, incy);
}
