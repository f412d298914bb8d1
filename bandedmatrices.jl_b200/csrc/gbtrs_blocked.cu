// gbtrs_blocked.cu -- wide-band solve with the factors of an interchange-free LU (ipiv = 1:n), panel-blocked.
//
// Replaces the column-by-column sweep of gbtrs_wide_kernel (gbtrs.cu) when LAPACK.gbtrs! (src/banded/linalg.jl:28) is
// called with factors whose pivot vector is the identity -- the case of every diagonally dominant system, e.g. the
// 2-D Laplacian of examples/finitedifference_2d.jl (BASELINE config C5).  The sweeps are still chains of n dependent
// steps, but per 16-column panel only a 16 x 16 triangle is sequential (one warp, shuffles); the (kl or kl+ku) x 16
// rectangle below / above it is one independent 16-term FMA chain per row, with the factor entries of the NEXT panel
// already in flight in registers while the triangle of the current one is being solved.
//
// Per element the operations and their order are exactly those of DGBTRS 'N' (SURVEY.md A.4): forward
// b[i] = fma(-b[j], L[i,j], b[i]) for j ascending; backward b[j] = b[j] / U[j,j] (true division), then
// b[i] = fma(-b[j], U[i,j], b[i]) for j descending -- so the solution is bit-identical to the reference path.
// One CTA per right-hand side; the active window of b lives in a shared-memory ring.
#include "common.cuh"

#define GB_THREADS 1024
#define GB_NB 16

// TMA bulk prefetch of a contiguous range into L2 (UBLKPF.L2): the range is widened to 16-byte alignment
__device__ __forceinline__ void gb_prefetch_l2_range(const double *p, int ndoubles)
{
    const unsigned long long a = (unsigned long long)p & ~15ull;
    const unsigned bytes = (unsigned)((((unsigned long long)(p + ndoubles) + 15ull) & ~15ull) - a);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(bytes) : "memory");
}

__global__ void gbtrs_count_interchanges(i64 n, const i64 *__restrict__ ipiv, int *__restrict__ out)
{
    int c = 0;
    for (i64 j = blockIdx.x * (i64)blockDim.x + threadIdx.x; j < n; j += (i64)gridDim.x * blockDim.x) c += (ipiv[j] != j + 1);
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

template <int KPL, int KPU>
__global__ void __launch_bounds__(GB_THREADS, 1)
gbtrs_wide_noswap(i64 n, int kl, int ku, const double *__restrict__ ab, i64 ldab, double *__restrict__ b, i64 ldb, int ring, int pfdist)
{
    extern __shared__ double rg[];
    __shared__ double xs[GB_NB];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, M = ring - 1;
    const int kv = kl + ku;
    constexpr int NB = GB_NB;
    double *x = b + (i64)blockIdx.x * ldb;
#ifdef GB_STATS
    long long cT[4] = {0, 0, 0, 0}, c0 = 0;
#define GB_T0() do { if ((tid & 31) == 0) c0 = clock64(); } while (0)
#define GB_T1(i) do { if ((tid & 31) == 0) cT[i] += clock64() - c0; } while (0)
#else
#define GB_T0() do { } while (0)
#define GB_T1(i) do { } while (0)
#endif
#define RG(row) rg[(int)(row) & M]
    // =========================================== forward: L y = b ===========================================
    if (kl > 0) {
        i64 hi = ((i64)kl + 2 * GB_THREADS < n) ? (i64)kl + 2 * GB_THREADS : n;  // rows [.., hi) resident
        for (i64 r = tid; r < hi; r += GB_THREADS) RG(r) = x[r];
        __syncthreads();
        auto window = [&](i64 j) {  // every 1024 columns: retire 1024 finished rows, pull 1024 new ones
            if (j >= GB_THREADS) x[j - GB_THREADS + tid] = RG(j - GB_THREADS + tid);
            __syncthreads();
            if (hi < n) {
                if (hi + tid < n) RG(hi + tid) = x[hi + tid];
                hi = (hi + GB_THREADS < n) ? hi + GB_THREADS : n;
            }
            __syncthreads();
        };
        // blocked panels: every row touched exists (J + NB + kl <= n)
        const i64 nblk = (n - kl >= NB) ? (n - kl) / NB : 0;
        // next panel's factor entries are pulled into L2 while this panel's triangle is solved: one 128-byte line per
        // thread (column jj = line / lpc holds kl contiguous doubles starting at L(J+NB, J+jj))
        auto prefetch_panel = [&](i64 J) {  // column J+jj from just below its diagonal (band row kv+1): triangle + rectangle
            if (tid >= 32 && tid < 32 + NB) {
                const int jj = tid - 32;
                gb_prefetch_l2_range(ab + (J + jj) * ldab + (kv + 1), kl);
            }
        };
        for (i64 p = 0; p < nblk; ++p) {
            const i64 J = p * NB;
            if ((J & (GB_THREADS - 1)) == 0) window(J);
            GB_T0();
            if (wid == 0) {  // 16 x 16 unit-lower triangle
                double Lt[NB];
#pragma unroll
                for (int jj = 0; jj < NB; ++jj) Lt[jj] = (lane < NB && jj < lane && lane - jj <= kl) ? ab[(kv + lane - jj) + (J + jj) * ldab] : 0.0;
                double xi = (lane < NB) ? RG(J + lane) : 0.0;
#pragma unroll
                for (int jj = 0; jj < NB - 1; ++jj) {
                    const double u = __shfl_sync(0xffffffffu, xi, jj);
                    if (lane > jj) xi = fma(-u, Lt[jj], xi);
                }
                if (lane < NB) { RG(J + lane) = xi; xs[lane] = xi; }
            } else if (pfdist > 0 && p + pfdist < nblk) {
                prefetch_panel(J + (i64)pfdist * NB);
            }
            GB_T1(0);
            __syncthreads();
            GB_T0();
#pragma unroll
            for (int k = 0; k < KPL; ++k) {
                const int t = tid + GB_THREADS * k;
                if (t < kl) {
                    double v[NB];
                    const double *pb = ab + J * ldab + (kv + NB + t);  // L(J+NB+t, J+jj) = pb[jj*(ldab-1)]
                    const int jmin = NB + t - kl;                      // in the band iff jj >= jmin
#pragma unroll
                    for (int jj = 0; jj < NB; ++jj) {
                        v[jj] = (jj >= jmin) ? *pb : 0.0;
                        pb += ldab - 1;
                    }
                    double acc = RG(J + NB + t);
#pragma unroll
                    for (int jj = 0; jj < NB; ++jj) acc = fma(-xs[jj], v[jj], acc);
                    RG(J + NB + t) = acc;
                }
            }
            GB_T1(1);
            __syncthreads();
        }
        // remaining columns one at a time
        for (i64 j = nblk * NB; j < n - 1; ++j) {
            if ((j & (GB_THREADS - 1)) == 0) window(j);
            const double t0 = -RG(j);
            __syncthreads();
#pragma unroll
            for (int k = 0; k < KPL; ++k) {
                const int i = 1 + tid + GB_THREADS * k;
                if (i <= kl && j + i < n) RG(j + i) = fma(t0, ab[(kv + i) + j * ldab], RG(j + i));
            }
            __syncthreads();
        }
        const i64 done = ((n - 2) >= 0) ? ((n - 2) & ~(i64)(GB_THREADS - 1)) : 0;
        for (i64 r = done + tid; r < n; r += GB_THREADS) x[r] = RG(r);
        __syncthreads();
    }
    // =========================================== backward: U x = y ===========================================
    {
        i64 lo = (n - ((i64)kv + 2 * GB_THREADS) > 0) ? n - ((i64)kv + 2 * GB_THREADS) : 0;
        for (i64 r = lo + tid; r < n; r += GB_THREADS) RG(r) = x[r];
        __syncthreads();
        auto window = [&](i64 j) {  // j = n-1-k with k a multiple of 1024
            const i64 k = n - 1 - j;
            if (k >= GB_THREADS) x[j + 1 + tid] = RG(j + 1 + tid);
            __syncthreads();
            if (lo > 0) {
                const i64 nlo = (lo - GB_THREADS > 0) ? lo - GB_THREADS : 0;
                if (nlo + tid < lo) RG(nlo + tid) = x[nlo + tid];
                lo = nlo;
            }
            __syncthreads();
        };
        // blocked panels [J, J+NB), from the bottom; every row above exists (J >= kv)
        const i64 nblk = (n - kv >= NB) ? (n - kv) / NB : 0;
        // column J+jj holds the kv entries above its diagonal contiguously: U(J+jj-kv .. J+jj-1, J+jj) = band rows 0 .. kv-1
        auto prefetch_panel = [&](i64 J) {  // band rows 0 .. kv of column J+jj: rectangle, triangle and diagonal
            if (tid >= 32 && tid < 32 + NB) {
                const int jj = tid - 32;
                gb_prefetch_l2_range(ab + (J + jj) * ldab, kv + 1);
            }
        };
        for (i64 p = 0; p < nblk; ++p) {
            const i64 J = n - (p + 1) * NB;
            if (((p * NB) & (GB_THREADS - 1)) == 0) window(J + NB - 1);
            GB_T0();
            if (wid == 0) {  // 16 x 16 upper triangle, columns descending; x / d correctly rounded (= true division)
                const int i = lane & 15;
                double Ut[NB];
                {
                    const double *pu = ab + J * ldab + (kv + i);  // U(J+i, J+jj) = pu[jj*(ldab-1)]
#pragma unroll
                    for (int jj = 0; jj < NB; ++jj) {
                        Ut[jj] = (jj > i) ? *pu : 0.0;   // jj - i <= 15 <= kv: always inside the band
                        pu += ldab - 1;
                    }
                }
                const double Ud = ab[kv + (J + i) * ldab];
                // One IEEE reciprocal per lane and panel, off the chain; each quotient on the chain is then two
                // Markstein corrections (5 dependent FMAs instead of a 131-cycle division sequence) -- see gb_div.
                const double rcp = 1.0 / Ud;
                const bool dsafe = gb_div_safe_divisor(Ud);
                double xi = RG(J + i);
#pragma unroll
                for (int jj = NB - 1; jj >= 0; --jj) {
                    if (i == jj) xi = gb_div(xi, Ud, rcp, dsafe);
                    const double q = __shfl_sync(0xffffffffu, xi, jj);
                    if (i < jj) xi = fma(-q, Ut[jj], xi);
                }
                if (lane < NB) { RG(J + lane) = xi; xs[lane] = xi; }
            } else if (pfdist > 0 && p + pfdist < nblk) {
                prefetch_panel(J - (i64)pfdist * NB);
            }
            GB_T1(2);
            __syncthreads();
            GB_T0();
#pragma unroll
            for (int k = 0; k < KPU; ++k) {
                const int t = tid + GB_THREADS * k;  // row J - 1 - t
                if (t < kv) {
                    double acc = RG(J - 1 - t);
                    const int jmax = kv - 1 - t;  // in the band iff jj <= jmax
#pragma unroll
                    for (int hf = 1; hf >= 0; --hf) {  // two halves of 8 columns (register budget), columns descending
                        double v[NB / 2];
                        const double *pb = ab + (J + hf * (NB / 2)) * ldab + (kv - 1 - t - hf * (NB / 2));  // U(J-1-t, J+jj) = pb[e*(ldab-1)]
#pragma unroll
                        for (int e = 0; e < NB / 2; ++e) {
                            v[e] = (hf * (NB / 2) + e <= jmax) ? *pb : 0.0;
                            pb += ldab - 1;
                        }
#pragma unroll
                        for (int e = NB / 2 - 1; e >= 0; --e) acc = fma(-xs[hf * (NB / 2) + e], v[e], acc);
                    }
                    RG(J - 1 - t) = acc;
                }
            }
            GB_T1(3);
            __syncthreads();
        }
        // remaining columns one at a time
        for (i64 j = n - 1 - nblk * NB; j >= 0; --j) {
            const i64 k = n - 1 - j;
            if ((k & (GB_THREADS - 1)) == 0) window(j);
            const double q = RG(j) / ab[kv + j * ldab];
            __syncthreads();
            if (tid == 0) RG(j) = q;
#pragma unroll
            for (int kk = 0; kk < KPU; ++kk) {
                const int i = 1 + tid + GB_THREADS * kk;
                if (i <= kv && j - i >= 0) RG(j - i) = fma(-q, ab[(kv - i) + j * ldab], RG(j - i));
            }
            __syncthreads();
        }
        const i64 lastk = ((n - 1) & ~(i64)(GB_THREADS - 1));
        const i64 top = n - 1 - lastk;
        for (i64 r = tid; r <= ((lastk >= GB_THREADS) ? top : n - 1); r += GB_THREADS) x[r] = RG(r);
    }
#ifdef GB_STATS
    if (blockIdx.x == 0 && (tid == 0 || tid == 32 || tid == 992))
        printf("[gbtrs_blocked] tid %d cycles: fwd tri %lld rect %lld | bwd tri %lld rect %lld  (n=%lld)\n", tid, cT[0], cT[1], cT[2], cT[3], (long long)n);
#endif
#undef RG
}

template <int KPL, int KPU>
static int launch_noswap(bmb200_ctx *h, i64 n, i64 kl, i64 ku, i64 nrhs, const double *dAB, i64 ldab, double *dB, i64 ldb)
{
    int ring = 4096;
    while (ring < kl + ku + 1 + 3 * GB_THREADS) ring <<= 1;
    const size_t smem = (size_t)ring * sizeof(double);
    BMB_CUDA(h, cudaFuncSetAttribute(gbtrs_wide_noswap<KPL, KPU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int pfdist = h->tune.gbtrs_pfdist_blocked;
    gbtrs_wide_noswap<KPL, KPU><<<(unsigned)nrhs, GB_THREADS, smem, h->stream>>>(n, (int)kl, (int)ku, dAB, ldab, dB, ldb, ring, pfdist);
    BMB_LAUNCH_CHECK(h);
    return 0;
}

int bmb_gbtrs_cluster(bmb200_ctx *h, i64 n, i64 kl, i64 ku, i64 nrhs, const double *dAB, i64 ldab, double *dB, i64 ldb);  // gbtrs_cluster.cu

// returns 1 when not applicable (the caller then runs the general kernel), 0 on success, <0 on error
// number of rows j with ipiv[j] != j+1 (synchronises); < 0: CUDA error code
int bmb_count_interchanges(bmb200_ctx *h, i64 n, const i64 *d_ipiv, int *count)
{
    int *cnt = h->d_info + 16;
    BMB_CUDA(h, cudaMemsetAsync(cnt, 0, sizeof(int), h->stream));
    gbtrs_count_interchanges<<<h->sm_count, 256, 0, h->stream>>>(n, d_ipiv, cnt);
    BMB_LAUNCH_CHECK(h);
    BMB_CUDA(h, cudaMemcpyAsync(count, cnt, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    BMB_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int bmb_gbtrs_blocked(bmb200_ctx *h, i64 n, i64 kl, i64 ku, i64 nrhs, const double *dAB, i64 ldab, const i64 *d_ipiv, double *dB, i64 ldb)
{
    if (h->tune.gbtrs_noblock) return 1;
    int hc = 0;
    const int rcc = bmb_count_interchanges(h, n, d_ipiv, &hc);
    if (rcc) return rcc;
    if (hc != 0) return 1;
    {   // one cluster per right-hand side, pipelined through distributed shared memory (gbtrs_cluster.cu)
        const int rc = bmb_gbtrs_cluster(h, n, kl, ku, nrhs, dAB, ldab, dB, ldb);
        if (rc != 1) return rc;
    }
    if (n < 4 * (kl + ku + GB_NB) || kl > GB_THREADS || kl + ku > 2 * GB_THREADS) return 1;
    return launch_noswap<1, 2>(h, n, kl, ku, nrhs, dAB, ldab, dB, ldb);
}

// ---- test hook (not part of the public ABI): counts inputs where gb_div differs from the IEEE quotient ----
__global__ void gb_divcheck_kernel(i64 n, const double *__restrict__ x, const double *__restrict__ d, unsigned long long *__restrict__ bad)
{
    for (i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x; t < n; t += (i64)gridDim.x * blockDim.x) {
        const double dv = d[t], xv = x[t];
        const double r = 1.0 / dv;
        const double q = gb_div(xv, dv, r, gb_div_safe_divisor(dv)), ref = xv / dv;
        if (__double_as_longlong(q) != __double_as_longlong(ref) && !(q != q && ref != ref)) atomicAdd(bad, 1ull);
    }
}
extern "C" int bmb200_internal_divcheck(bmb200_handle_t h, int64_t n, const double *dx, const double *dd, unsigned long long *dbad)
{
    if (!h) return -1;
    DeviceGuard g(h->device);
    gb_divcheck_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(n, dx, dd, dbad);
    BMB_LAUNCH_CHECK(h);
    return 0;
}
