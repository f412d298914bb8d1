// tb.cu -- triangular band solve / multiply on the device: tbsv! and tbmv! of the reference (src/blas.jl:71-141), reached
// from ldiv!/lmul! of UpperTriangular / LowerTriangular{<:BandedMatrix} (src/tribanded.jl:47-84).  SURVEY.md 8(f) rank 2:
// the same sweeps as the back substitution inside gbtrs, exposed on BLAS triangular-band storage
//   'U': T[i,j] at a[(k + i - j) + j*lda]      'L': T[i,j] at a[(i - j) + j*lda]        (0-based, i - j within the band)
// trans = 'N' and 'T'/'C' (the reference reaches 'T' through row-major layouts, src/tribanded.jl:86-96).  incx must be 1.
//   * bmb200_dtbsv: the cluster pipeline of gbtrs_cluster.cu, one sweep (OpenBLAS tbsv_{U,L}: true division by the
//     diagonal unless unit, then fma(-x_j, T[i,j], x_i)) => bit-identical to the reference CPU path.
//   * bmb200_dtbmv: one thread per row, d_i*x_i (x_i if unit) then fma(x_j, T[i,j], .) over j ascending ('U') or
//     descending ('L') -- the per-row order of OpenBLAS' column sweep -- into a scratch vector, then copied back.
#include "common.cuh"

int bmb_tri_solve_via_gbtrs(bmb200_ctx *h, int up, int tr, int unit, i64 n, i64 k, i64 nrhs, const double *dA, i64 lda, double *dB, i64 ldb);  // pb.cu
int bmb_tri_solve_transposed_wide(bmb200_ctx *h, int up, int unit, i64 n, i64 k, i64 nrhs, const double *dA, i64 lda, double *dB, i64 ldb);  // pb.cu
int bmb_cluster_solve(bmb200_ctx *h, int mode, i64 n, i64 kl, i64 ku, i64 nrhs, const double *dAB, i64 ldab, double *dB, i64 ldb);  // gbtrs_cluster.cu

__global__ void __launch_bounds__(256)
tbmv_rows(i64 n, int k, int up, int unit, const double *__restrict__ a, i64 lda, const double *__restrict__ x, double *__restrict__ y)
{
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        double acc = x[i];
        if (up) {
            if (!unit) acc = __dmul_rn(acc, a[k + i * lda]);
            const i64 j1 = (i + k < n - 1) ? i + k : n - 1;
            const double *p = a + (k - 1) + (i + 1) * lda;  // T[i, i+1]; next column: + lda - 1
            for (i64 j = i + 1; j <= j1; ++j, p += lda - 1) acc = fma(x[j], *p, acc);
        } else {
            if (!unit) acc = __dmul_rn(acc, a[i * lda]);
            const i64 j0 = (i - k > 0) ? i - k : 0;
            const double *p = a + 1 + (i - 1) * lda;        // T[i, i-1]; previous column: - lda + 1
            for (i64 j = i - 1; j >= j0; --j, p -= lda - 1) acc = fma(x[j], *p, acc);
        }
        y[i] = acc;
    }
}

// Wide bands: a warp owns 32 consecutive rows (lane = row) and sweeps the columns they reach; for a fixed column the 32 lanes
// read 32 consecutive band entries (one coalesced 256-byte request) and x[j] is a broadcast.  Each row still accumulates in
// OpenBLAS' order (j ascending for 'U', descending for 'L').  Columns go in chunks of 8; the next chunk's loads are issued
// before the current chunk's FMA chain (register double buffer), and chunks that lie inside the band for all 32 rows -- all
// but ~40 columns per row block -- run without any predicate.
template <bool UP>
__device__ __forceinline__ bool tbmv_load8(const double *__restrict__ p, i64 st, const double *__restrict__ x, i64 j0, i64 lo_all, i64 hi_all,
                                           i64 jlo, i64 jhi, i64 i, int k, bool live, double (&v)[8], double (&xv)[8])
{
    // chunk columns: UP j0 .. j0+7, else j0 .. j0-7;  [lo_all, hi_all] = columns every one of the 32 rows reaches
    const i64 ja = UP ? j0 : j0 - 7, jb = UP ? j0 + 7 : j0;
    if (ja >= lo_all && jb <= hi_all) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const i64 j = UP ? j0 + e : j0 - e;
            v[e] = ld_stream(p + j * st);
            xv[e] = x[j];
        }
        return true;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const i64 j = UP ? j0 + e : j0 - e;
        const bool inr = j >= jlo && j <= jhi;
        const bool ok = live && inr && (UP ? (j > i && j - i <= k) : (j < i && i - j <= k));
        v[e] = ok ? ld_stream(p + j * st) : 0.0;
        xv[e] = inr ? x[j] : 0.0;
    }
    return false;
}
template <bool UP>
__device__ __forceinline__ double tbmv_fma8(double acc, bool full, i64 j0, i64 jlo, i64 jhi, i64 i, int k, bool live, const double (&v)[8],
                                            const double (&xv)[8])
{
    if (full) {
#pragma unroll
        for (int e = 0; e < 8; ++e) acc = fma(xv[e], v[e], acc);
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const i64 j = UP ? j0 + e : j0 - e;
            if (live && j >= jlo && j <= jhi && (UP ? (j > i && j - i <= k) : (j < i && i - j <= k))) acc = fma(xv[e], v[e], acc);
        }
    }
    return acc;
}

template <bool UP>
__global__ void __launch_bounds__(256)
tbmv_sweep(i64 n, int k, int unit, const double *__restrict__ a, i64 lda, const double *__restrict__ x, double *__restrict__ y)
{
    const int lane = threadIdx.x & 31;
    const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    const i64 st = lda - 1;
    for (i64 rb = warp * 32; rb < n; rb += nwarps * 32) {
        const i64 i = rb + lane;
        const bool live = i < n;
        double acc = live ? x[i] : 0.0;
        if (live && !unit) acc = __dmul_rn(acc, a[(UP ? k : 0) + i * lda]);
        // T[i,j] = p[j*(lda-1)] with p = a + (UP ? k : 0) + i
        const double *p = a + (UP ? k : 0) + i;
        // columns reached by the block [jlo, jhi]; by all 32 rows (only if all 32 rows exist) [lo_all, hi_all]
        const i64 jlo = UP ? rb + 1 : ((rb - k > 0) ? rb - k : 0);
        const i64 jhi = UP ? ((rb + 31 + k < n - 1) ? rb + 31 + k : n - 1) : rb + 30;
        const bool all = rb + 31 < n;
        const i64 lo_all = all ? (UP ? rb + 32 : ((rb + 31 - k > 0) ? rb + 31 - k : 0)) : 1;
        const i64 hi_all = all ? (UP ? ((rb + k < n - 1) ? rb + k : n - 1) : rb - 1) : 0;
        if (jlo > jhi) { if (live) y[i] = acc; continue; }
        const i64 nch = (jhi - jlo) / 8 + 1;
        double va[8], xa[8], vb[8], xb[8];
        bool fa, fb = false;
        i64 j0 = UP ? jlo : jhi;
        fa = tbmv_load8<UP>(p, st, x, j0, lo_all, hi_all, jlo, jhi, i, k, live, va, xa);
        for (i64 c = 0; c < nch; c += 2) {
            const i64 j1 = UP ? j0 + 8 : j0 - 8;
            if (c + 1 < nch) fb = tbmv_load8<UP>(p, st, x, j1, lo_all, hi_all, jlo, jhi, i, k, live, vb, xb);
            acc = tbmv_fma8<UP>(acc, fa, j0, jlo, jhi, i, k, live, va, xa);
            const i64 j2 = UP ? j1 + 8 : j1 - 8;
            if (c + 2 < nch) fa = tbmv_load8<UP>(p, st, x, j2, lo_all, hi_all, jlo, jhi, i, k, live, va, xa);
            if (c + 1 < nch) acc = tbmv_fma8<UP>(acc, fb, j1, jlo, jhi, i, k, live, vb, xb);
            j0 = j2;
        }
        if (live) y[i] = acc;
    }
}

// ---- trans = 'T': op(T) = T^T.  Every column of the band array is now a ROW of the operator: a dot product of contiguous
// band entries with a window of x.  OpenBLAS uses its SIMD dot kernel here (summation order unspecified), so these are compared
// to the oracle at 1e-13 like the other transposed paths, not bit for bit. ----
// tbmv 'T': one warp per column, lanes stride the contiguous column, shuffle reduction.
__global__ void __launch_bounds__(256)
tbmv_t_cols(i64 n, int k, int up, int unit, const double *__restrict__ a, i64 lda, const double *__restrict__ x, double *__restrict__ y)
{
    const int lane = threadIdx.x & 31;
    const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    for (i64 j = warp; j < n; j += nwarps) {
        // 'U': rows j-k..j-1 live at a[k-(j-i) + j*lda]; 'L': rows j+1..j+k at a[(i-j) + j*lda]
        const i64 i0 = up ? ((j - k > 0) ? j - k : 0) : j + 1, i1 = up ? j - 1 : ((j + k < n - 1) ? j + k : n - 1);
        const double *col = a + j * lda + (up ? k - j : -j);  // T[i,j] = col[i]
        double acc = 0.0;
        for (i64 i = i0 + lane; i <= i1; i += 32) acc = fma(ld_stream(col + i), x[i], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) y[j] = (unit ? x[j] : __dmul_rn(x[j], a[(up ? k : 0) + j * lda])) + acc;
    }
}

// tbsv 'T': a chain of n dependent dot products, one warp: x_j = (x_j - sum_i T[i,j] x_i) / T[j,j], j ascending for 'U'
// (T^T is lower triangular), descending for 'L'.  The solved entries the next columns need stay in a shared-memory ring; the
// band entries of the next column are loaded (they do not depend on x) before the current column's reduction.
template <int KPL>  // band entries per lane: 32*KPL >= k
__global__ void __launch_bounds__(32)
tbsv_t_chain(i64 n, int k, int up, int unit, const double *__restrict__ a, i64 lda, double *__restrict__ x0, i64 ldx, int ring)
{
    extern __shared__ double xr[];  // ring of solved entries, indexed by matrix row & (ring-1)
    double *__restrict__ x = x0 + (i64)blockIdx.x * ldx;  // one chain block per right-hand side
    const int lane = threadIdx.x, M = ring - 1;
    auto colptr = [&](i64 j) { return a + j * lda + (up ? k : 0); };  // diagonal entry of column j
    auto load_col = [&](i64 j, double (&v)[KPL]) {
        // entry e = 1..k of column j pairs with row j-e ('U') / j+e ('L'); lane handles e = 1 + lane + 32*q
        const double *d = colptr(j);
#pragma unroll
        for (int q = 0; q < KPL; ++q) {
            const int e = 1 + lane + 32 * q;
            const i64 r = up ? j - e : j + e;
            v[q] = (e <= k && r >= 0 && r < n) ? ld_stream(up ? d - e : d + e) : 0.0;
        }
    };
    double v[KPL], vn[KPL];
    i64 j = up ? 0 : n - 1;
    const i64 step = up ? 1 : -1;
    load_col(j, v);
    for (i64 c = 0; c < n; ++c, j += step) {
        const i64 jn = j + step;
        if (c + 1 < n) load_col(jn, vn);
        const double dj = unit ? 1.0 : colptr(j)[0];
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < KPL; ++q) {
            const int e = 1 + lane + 32 * q;
            const i64 r = up ? j - e : j + e;
            if (e <= k && r >= 0 && r < n) acc = fma(v[q], xr[(int)(r & M)], acc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        double xj = x[j] - acc;
        if (!unit) xj = xj / dj;
        if (lane == 0) { xr[(int)(j & M)] = xj; x[j] = xj; }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < KPL; ++q) v[q] = vn[q];
    }
}

// ---- symmetric band matvec: y <- alpha*S*x + beta*y, S symmetric with only its 'U' or 'L' triangle stored in triangular-band
// storage (sbmv!, src/blas.jl:36-66; mul! of Symmetric{<:BandedMatrix}, src/symbanded/symbanded.jl:72-93).  SURVEY 8(f) rank 3.
// Row i needs the stored triangle twice: S[i,j] on the stored side of the diagonal is a strided walk through OTHER columns
// (coalesced over the 32 rows of a warp, exactly the tbmv sweep), and on the other side it is the lane's OWN column, contiguous
// per lane.  OpenBLAS' dsbmv mixes an axpy and a SIMD dot per column (summation order unspecified): compared at 1e-13. ----
template <bool UP>
__global__ void __launch_bounds__(256)
sbmv_sweep(i64 n, int k, double alpha, const double *__restrict__ a, i64 lda, const double *__restrict__ x, double beta, double *__restrict__ y)
{
    const int lane = threadIdx.x & 31;
    const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    const i64 st = lda - 1;
    for (i64 rb = warp * 32; rb < n; rb += nwarps * 32) {
        const i64 i = rb + lane;
        const bool live = i < n;
        if (alpha == 0.0) {  // DSBMV returns after the beta scaling: A and x are not read
            if (live) y[i] = (beta == 0.0) ? 0.0 : __dmul_rn(beta, y[i]);
            continue;
        }
        // (1) own column: diagonal + the k entries on the stored side, contiguous: 'U' rows i-k..i-1 above, 'L' rows i+1..i+k below
        double acc = 0.0;
        if (live) {
            const double *col = a + i * lda + (UP ? k : 0);  // diagonal entry; S[i-d, i] = col[-d] ('U'), S[i+d, i] = col[+d] ('L')
            const i64 dmax = UP ? ((i < k) ? i : k) : ((n - 1 - i < k) ? n - 1 - i : k);
            const double *xp = x + i;
            double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;  // four chains: a 1024-term dependent FMA chain would be latency bound
            i64 d = 1;
#pragma unroll 2
            for (; d + 3 <= dmax; d += 4) {
                c0 = fma(UP ? col[-d] : col[d], UP ? xp[-d] : xp[d], c0);
                c1 = fma(UP ? col[-d - 1] : col[d + 1], UP ? xp[-d - 1] : xp[d + 1], c1);
                c2 = fma(UP ? col[-d - 2] : col[d + 2], UP ? xp[-d - 2] : xp[d + 2], c2);
                c3 = fma(UP ? col[-d - 3] : col[d + 3], UP ? xp[-d - 3] : xp[d + 3], c3);
            }
            for (; d <= dmax; ++d) c0 = fma(UP ? col[-d] : col[d], UP ? xp[-d] : xp[d], c0);
            acc = fma(col[0], xp[0], (c0 + c1) + (c2 + c3));
        }
        // (2) the other side of the diagonal: S[i,j] = stored entry (row i, column j), j > i ('U') or j < i ('L'): column sweep
        const double *p = a + (UP ? k : 0) + i;  // stored T[i,j] = p[j*(lda-1)]
        const i64 jlo = UP ? rb + 1 : ((rb - k > 0) ? rb - k : 0);
        const i64 jhi = UP ? ((rb + 31 + k < n - 1) ? rb + 31 + k : n - 1) : rb + 30;
        const bool all = rb + 31 < n;
        const i64 lo_all = all ? (UP ? rb + 32 : ((rb + 31 - k > 0) ? rb + 31 - k : 0)) : 1;
        const i64 hi_all = all ? (UP ? ((rb + k < n - 1) ? rb + k : n - 1) : rb - 1) : 0;
        if (jlo <= jhi) {
            const i64 nch = (jhi - jlo) / 8 + 1;
            double va[8], xa[8], vb[8], xb[8];
            bool fa, fb = false;
            i64 j0 = UP ? jlo : jhi;
            fa = tbmv_load8<UP>(p, st, x, j0, lo_all, hi_all, jlo, jhi, i, k, live, va, xa);
            for (i64 c = 0; c < nch; c += 2) {
                const i64 j1 = UP ? j0 + 8 : j0 - 8;
                if (c + 1 < nch) fb = tbmv_load8<UP>(p, st, x, j1, lo_all, hi_all, jlo, jhi, i, k, live, vb, xb);
                acc = tbmv_fma8<UP>(acc, fa, j0, jlo, jhi, i, k, live, va, xa);
                const i64 j2 = UP ? j1 + 8 : j1 - 8;
                if (c + 2 < nch) fa = tbmv_load8<UP>(p, st, x, j2, lo_all, hi_all, jlo, jhi, i, k, live, va, xa);
                if (c + 1 < nch) acc = tbmv_fma8<UP>(acc, fb, j1, jlo, jhi, i, k, live, vb, xb);
                j0 = j2;
            }
        }
        if (live) y[i] = (beta == 0.0) ? __dmul_rn(alpha, acc) : fma(alpha, acc, __dmul_rn(beta, y[i]));
    }
}

// narrow bands: one thread per row; both uses of a stored entry fall into the same or the neighbouring thread's cache lines
template <bool UP>
__global__ void __launch_bounds__(256)
sbmv_rows(i64 n, int k, double alpha, const double *__restrict__ a, i64 lda, const double *__restrict__ x, double beta, double *__restrict__ y)
{
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        if (alpha == 0.0) {
            y[i] = (beta == 0.0) ? 0.0 : __dmul_rn(beta, y[i]);
            continue;
        }
        const double *col = a + i * lda + (UP ? k : 0);  // diagonal of the own column
        double acc = col[0] * x[i];
        for (int d = 1; d <= k; ++d) {
            const i64 js = UP ? i - d : i + d;  // stored side: own column, entry col[-d] / col[+d]
            const i64 jo = UP ? i + d : i - d;  // other side: entry (row i) of column jo
            if (js >= 0 && js < n) acc = fma(UP ? col[-d] : col[d], x[js], acc);
            if (jo >= 0 && jo < n) acc = fma(a[(UP ? k - d : d) + jo * lda], x[jo], acc);
        }
        y[i] = (beta == 0.0) ? __dmul_rn(alpha, acc) : fma(alpha, acc, __dmul_rn(beta, y[i]));
    }
}

extern "C" int bmb200_dsbmv(bmb200_handle_t h, char uplo, int64_t n, int64_t k, double alpha, const double *dA, int64_t lda, const double *dx,
                            int64_t incx, double beta, double *dy, int64_t incy)
{
    if (!h) return -1;
    const int up = (uplo == 'U' || uplo == 'u');
    if (!up && !(uplo == 'L' || uplo == 'l')) return -2;
    if (n < 0) return -3;
    if (k < 0 || k >= ((int64_t)1 << 30)) return -4;
    if (lda < k + 1) return -7;
    if (incx != 1) return -9;
    if (incy != 1) return -12;
    if (n == 0) return 0;
    if (!dA || !dx || !dy) return -6;
    if (dx == dy) return -8;  // the reference un-aliases x and y before the call (symbanded.jl:77-83)
    DeviceGuard g(h->device);
    const int kthr = h->tune.sbmv_rows_k;
    if (k < kthr) {
        const i64 blocks = imin64(cdiv64(n, 256), (i64)h->sm_count * 16);
        if (up) sbmv_rows<true><<<(unsigned)blocks, 256, 0, h->stream>>>(n, (int)k, alpha, dA, lda, dx, beta, dy);
        else sbmv_rows<false><<<(unsigned)blocks, 256, 0, h->stream>>>(n, (int)k, alpha, dA, lda, dx, beta, dy);
    } else {
        const i64 blocks = imin64(cdiv64(n, 32 * 8), (i64)h->sm_count * 8);
        if (up) sbmv_sweep<true><<<(unsigned)blocks, 256, 0, h->stream>>>(n, (int)k, alpha, dA, lda, dx, beta, dy);
        else sbmv_sweep<false><<<(unsigned)blocks, 256, 0, h->stream>>>(n, (int)k, alpha, dA, lda, dx, beta, dy);
    }
    BMB_LAUNCH_CHECK(h);
    return 0;
}

static int tb_check(char uplo, char trans, char diag, int64_t n, int64_t k, int64_t lda, int64_t incx, int &up, int &unit)
{
    up = (uplo == 'U' || uplo == 'u');
    unit = (diag == 'U' || diag == 'u');
    if (!up && !(uplo == 'L' || uplo == 'l')) return -2;
    if (!(trans == 'N' || trans == 'n' || trans == 'T' || trans == 't' || trans == 'C' || trans == 'c')) return -3;
    if (!unit && !(diag == 'N' || diag == 'n')) return -4;
    if (n < 0) return -5;
    if (k < 0) return -6;
    if (lda < k + 1) return -8;
    if (incx != 1) return -10;
    return 0;
}

// tbsv 'T' for nrhs right-hand sides (columns of dB, stride ldb): one chain block each.  Also the U^T / L^T sweep of dpbtrs (pb.cu).
int bmb_tbsv_t_multi(bmb200_ctx *h, int up, int unit, i64 n, i64 k, i64 nrhs, const double *dA, i64 lda, double *dB, i64 ldb)
{
    if (k > 32 * 32) {
        snprintf(h->err, sizeof(h->err), "dtbsv 'T': band width %lld > 1024 is not supported", (long long)k);
        return BMB200_ERR_CUDA;
    }
    int ring = 64;
    while (ring < k + 2) ring <<= 1;
    const size_t smem = (size_t)ring * sizeof(double);
#define TB_T_LAUNCH(KPL) tbsv_t_chain<KPL><<<(unsigned)nrhs, 32, smem, h->stream>>>(n, (int)k, up, unit, dA, lda, dB, ldb, ring)
    if (k <= 32) TB_T_LAUNCH(1);
    else if (k <= 64) TB_T_LAUNCH(2);
    else if (k <= 128) TB_T_LAUNCH(4);
    else if (k <= 256) TB_T_LAUNCH(8);
    else if (k <= 512) TB_T_LAUNCH(16);
    else TB_T_LAUNCH(32);
#undef TB_T_LAUNCH
    BMB_LAUNCH_CHECK(h);
    return 0;
}

extern "C" int bmb200_dtbsv(bmb200_handle_t h, char uplo, char trans, char diag, int64_t n, int64_t k, const double *dA, int64_t lda,
                            double *dx, int64_t incx)
{
    if (!h) return -1;
    int up, unit;
    const int rc0 = tb_check(uplo, trans, diag, n, k, lda, incx, up, unit);
    if (rc0) return rc0;
    if (n == 0) return 0;
    if (!dA || !dx) return -7;
    DeviceGuard g(h->device);
    if (!(trans == 'N' || trans == 'n')) {
        // non-unit transposed solves run as column sweeps on a transposed (wide) or reversed / transposed (narrow) copy of the
        // factor: the single-warp chain of dot products costs 250 ns (k = 4) to 2 us (k = 1024) per column
        if (k <= 63 && n > 1) return bmb_tri_solve_via_gbtrs(h, up, 1, unit, n, k, 1, dA, lda, dx, n);
        if (n > 1) {
            const int rc = bmb_tri_solve_transposed_wide(h, up, unit, n, k, 1, dA, lda, dx, n);
            if (rc != 1) return rc;
        }
        return bmb_tbsv_t_multi(h, up, unit, n, k, 1, dA, lda, dx, n > 1 ? n : 1);
    }
    // narrow bands, non-unit diagonal: the multi-RHS back substitution of bmb200_dgbtrs (register-window kernels, ~45 ns per
    // column) instead of the cluster pipeline, which is built for wide bands (~250 ns per column at k = 4); pb.cu
    if (k <= 63 && n > 1) return bmb_tri_solve_via_gbtrs(h, up, 0, unit, n, k, 1, dA, lda, dx, n);
    // 'U': diagonal in row k of the band array, reach k above it (mode 0 with kl = 0 divides, mode 1 does not);
    // 'L': diagonal in row 0, reach k below it (mode 2 unit, mode 3 dividing)
    const int mode = up ? (unit ? 1 : 0) : (unit ? 2 : 3);
    const int rc = bmb_cluster_solve(h, mode, n, up ? 0 : k, up ? k : 0, 1, dA, lda, dx, n > 1 ? n : 1);
    if (rc == 1) {
        snprintf(h->err, sizeof(h->err), "dtbsv: band width %lld is not supported by the cluster pipeline on this device", (long long)k);
        return BMB200_ERR_CUDA;
    }
    return rc;
}

extern "C" int bmb200_dtbmv(bmb200_handle_t h, char uplo, char trans, char diag, int64_t n, int64_t k, const double *dA, int64_t lda,
                            double *dx, int64_t incx)
{
    if (!h) return -1;
    int up, unit;
    const int rc0 = tb_check(uplo, trans, diag, n, k, lda, incx, up, unit);
    if (rc0) return rc0;
    if (n == 0) return 0;
    if (!dA || !dx) return -7;
    if (k >= ((int64_t)1 << 30)) return -6;
    DeviceGuard g(h->device);
    if (bmb_ensure_scratch(h, (size_t)n * sizeof(double)) != 0) return BMB200_ERR_CUDA;
    double *y = (double *)h->scratch;
    if (!(trans == 'N' || trans == 'n')) {
        const i64 blocks = imin64(cdiv64(n, 8), (i64)h->sm_count * 8);
        tbmv_t_cols<<<(unsigned)blocks, 256, 0, h->stream>>>(n, (int)k, up, unit, dA, lda, dx, y);
    } else if (k >= 16) {
        const i64 blocks = imin64(cdiv64(n, 32 * 8), (i64)h->sm_count * 8);
        if (up) tbmv_sweep<true><<<(unsigned)blocks, 256, 0, h->stream>>>(n, (int)k, unit, dA, lda, dx, y);
        else tbmv_sweep<false><<<(unsigned)blocks, 256, 0, h->stream>>>(n, (int)k, unit, dA, lda, dx, y);
    } else {
        const i64 blocks = imin64(cdiv64(n, 256), (i64)h->sm_count * 16);
        tbmv_rows<<<(unsigned)blocks, 256, 0, h->stream>>>(n, (int)k, up, unit, dA, lda, dx, y);
    }
    BMB_LAUNCH_CHECK(h);
    BMB_CUDA(h, cudaMemcpyAsync(dx, y, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    return 0;
}
