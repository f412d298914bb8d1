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

__device__ __forceinline__ void gb_prefetch_l2(const double *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__global__ void gbtrs_count_interchanges(i64 n, const i64 *__restrict__ ipiv, int *__restrict__ out)
{
    int c = 0;
    for (i64 j = blockIdx.x * (i64)blockDim.x + threadIdx.x; j < n; j += (i64)gridDim.x * blockDim.x) c += (ipiv[j] != j + 1);
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

template <int KPL, int KPU>
__global__ void __launch_bounds__(GB_THREADS, 1)
gbtrs_wide_noswap(i64 n, int kl, int ku, const double *__restrict__ ab, i64 ldab, double *__restrict__ b, i64 ldb, int ring)
{
    extern __shared__ double rg[];
    __shared__ double xs[GB_NB];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, M = ring - 1;
    const int kv = kl + ku;
    constexpr int NB = GB_NB;
    double *x = b + (i64)blockIdx.x * ldb;
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
        const int lpc = (kl * 8 + 127) / 128 + 1;  // lines per column, alignment slack included
        auto prefetch_panel = [&](i64 J) {
            for (int ln = tid; ln < NB * lpc; ln += GB_THREADS) {
                const int jj = ln / lpc, off = (ln - jj * lpc) * 16;
                if (off < kl + 16) gb_prefetch_l2(ab + (J + jj) * ldab + (kv + NB - jj) + (off < kl ? off : kl - 1));
            }
        };
        for (i64 p = 0; p < nblk; ++p) {
            const i64 J = p * NB;
            if ((J & (GB_THREADS - 1)) == 0) window(J);
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
            } else if (p + 1 < nblk) {
                prefetch_panel(J + NB);
            }
            __syncthreads();
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
        const int lpc = (kv * 8 + 127) / 128 + 1;
        auto prefetch_panel = [&](i64 J) {
            for (int ln = tid; ln < NB * lpc; ln += GB_THREADS) {
                const int jj = ln / lpc, off = (ln - jj * lpc) * 16;
                if (off < kv + 16) gb_prefetch_l2(ab + (J + jj) * ldab + (off < kv ? off : kv - 1));
            }
        };
        for (i64 p = 0; p < nblk; ++p) {
            const i64 J = n - (p + 1) * NB;
            if (((p * NB) & (GB_THREADS - 1)) == 0) window(J + NB - 1);
            if (wid == 0) {  // 16 x 16 upper triangle, columns descending; true division by the diagonal
                double Ut[NB];
#pragma unroll
                for (int jj = 0; jj < NB; ++jj) Ut[jj] = (lane < NB && jj > lane && jj - lane <= kv) ? ab[(kv - (jj - lane)) + (J + jj) * ldab] : 0.0;
                const double Ud = (lane < NB) ? ab[kv + (J + lane) * ldab] : 1.0;
                double xi = (lane < NB) ? RG(J + lane) : 0.0;
#pragma unroll
                for (int jj = NB - 1; jj >= 0; --jj) {
                    if (lane == jj) xi = xi / Ud;
                    const double q = __shfl_sync(0xffffffffu, xi, jj);
                    if (lane < jj) xi = fma(-q, Ut[jj], xi);
                }
                if (lane < NB) { RG(J + lane) = xi; xs[lane] = xi; }
            } else if (p + 1 < nblk) {
                prefetch_panel(J - NB);
            }
            __syncthreads();
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
#undef RG
}

template <int KPL, int KPU>
static int launch_noswap(bmb200_ctx *h, i64 n, i64 kl, i64 ku, i64 nrhs, const double *dAB, i64 ldab, double *dB, i64 ldb)
{
    int ring = 4096;
    while (ring < kl + ku + 1 + 3 * GB_THREADS) ring <<= 1;
    const size_t smem = (size_t)ring * sizeof(double);
    BMB_CUDA(h, cudaFuncSetAttribute(gbtrs_wide_noswap<KPL, KPU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gbtrs_wide_noswap<KPL, KPU><<<(unsigned)nrhs, GB_THREADS, smem, h->stream>>>(n, (int)kl, (int)ku, dAB, ldab, dB, ldb, ring);
    BMB_LAUNCH_CHECK(h);
    return 0;
}

// returns 1 when not applicable (the caller then runs the general kernel), 0 on success, <0 on error
int bmb_gbtrs_blocked(bmb200_ctx *h, i64 n, i64 kl, i64 ku, i64 nrhs, const double *dAB, i64 ldab, const i64 *d_ipiv, double *dB, i64 ldb)
{
    static const bool off = getenv("BMB200_GBTRS_NOBLOCK") != nullptr;
    if (off || n < 4 * (kl + ku + GB_NB) || kl > GB_THREADS || kl + ku > 2 * GB_THREADS) return 1;
    int *cnt = h->d_info + 16;
    BMB_CUDA(h, cudaMemsetAsync(cnt, 0, sizeof(int), h->stream));
    gbtrs_count_interchanges<<<h->sm_count, 256, 0, h->stream>>>(n, d_ipiv, cnt);
    BMB_LAUNCH_CHECK(h);
    int hc = 0;
    BMB_CUDA(h, cudaMemcpyAsync(&hc, cnt, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    BMB_CUDA(h, cudaStreamSynchronize(h->stream));
    if (hc != 0) return 1;
    return launch_noswap<1, 2>(h, n, kl, ku, nrhs, dAB, ldab, dB, ldb);
}
