// gbtrs.cu -- multi-RHS solve with the band LU factors on sm_100a.
//
// Replaces dgbtrs_ (LAPACK.gbtrs! at src/banded/linalg.jl:28 'N', :46 'T', :62 'C').
// 'N' keeps the reference operation order: forward sweep = row swap then one FMA
// per (row, rhs) with t = -b[j]; backward sweep = DTBSV upper/no-trans/non-unit with a TRUE division
// by the diagonal and one FMA per term -- the solution is bit-identical to OpenBLAS' dgbtrs.
// Both sweeps are chains of n dependent steps (the row swaps are interleaved with the elimination
// because L is stored un-permuted), so parallelism comes from the right-hand sides: a warp owns NR
// RHS columns whose active window lives in a shared-memory ring; lanes span the band; the L / U
// column of the NEXT steps is prefetched into registers so that no global latency sits on the chain.
#include "common.cuh"

#define GBTRS_WARPS 4

int bmb_gbtrs_slot(bmb200_ctx *h, i64 n, i64 kl, i64 ku, i64 nrhs, const double *dAB, i64 ldab, const i64 *d_ipiv,
                   double *dB, i64 ldb);  // gbtrs_slot.cu
int bmb_gbtrs_shfl(bmb200_ctx *h, i64 n, i64 kl, i64 ku, i64 nrhs, const double *dAB, i64 ldab, const i64 *d_ipiv,
                   double *dB, i64 ldb);
int bmb_gbtrs_blocked(bmb200_ctx *h, i64 n, i64 kl, i64 ku, i64 nrhs, const double *dAB, i64 ldab, const i64 *d_ipiv, double *dB, i64 ldb);  // gbtrs_shfl.cu

// values per lane KPL = ceil(band/32); SB steps are prefetched together.
template <int NR, int KPL, int SB>
__global__ void __launch_bounds__(GBTRS_WARPS * 32)
gbtrs_n_kernel(i64 n, int kl, int ku, i64 nrhs, const double *__restrict__ ab, i64 ldab,
               const i64 *__restrict__ ipiv, double *__restrict__ b, i64 ldb, int ring, int skip_u)
{
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const i64 tile = (i64)blockIdx.x * GBTRS_WARPS + wid;
    const i64 c0 = tile * NR;
    if (c0 >= nrhs) return;
    const int kv = kl + ku, M = ring - 1;
    double *rg = sm + (size_t)wid * NR * ring;  // rg[q*ring + (row & M)]
#define RG(q, row) rg[(q) * ring + ((int)(row) & M)]
    const int nq = (int)((nrhs - c0 < NR) ? (nrhs - c0) : NR);

    // ---------------- forward: L y = P b ----------------
    if (kl > 0) {
        // rows [0, hi) are resident in the ring
        i64 hi = (kl + 64 < n) ? kl + 64 : n;
        for (int q = 0; q < nq; ++q)
            for (i64 r = lane; r < hi; r += 32) RG(q, r) = b[r + (c0 + q) * ldb];
        __syncwarp();
        double L[SB][KPL], Ln[SB][KPL];
        i64 pv = 0, pvn = 0;
        auto loadL = [&](i64 j0, double (&dst)[SB][KPL], i64 &pdst) {
#pragma unroll
            for (int s = 0; s < SB; ++s) {
                const i64 j = j0 + s;
#pragma unroll
                for (int k = 0; k < KPL; ++k) {
                    const int i = 1 + lane + 32 * k;
                    dst[s][k] = (j < n - 1 && i <= kl && j + i < n) ? ab[(kv + i) + j * ldab] : 0.0;
                }
            }
            pdst = (lane < SB && j0 + lane < n - 1) ? ipiv[j0 + lane] - 1 : 0;
        };
        loadL(0, L, pv);
        for (i64 j0 = 0; j0 < n - 1; j0 += SB) {
            loadL(j0 + SB, Ln, pvn);
#pragma unroll
            for (int s = 0; s < SB; ++s) {
                const i64 j = j0 + s;
                if (j < n - 1) {
                    // keep the window ahead of the band: every 32 steps pull 32 more rows, retire 32 old ones
                    if ((j & 31) == 0) {
                        if (j >= 32)
                            for (int q = 0; q < nq; ++q) b[(j - 32 + lane) + (c0 + q) * ldb] = RG(q, j - 32 + lane);
                        __syncwarp();  // retired rows are read before their ring slots are reused
                        const i64 r = hi + lane;
                        if (hi < n) {
                            for (int q = 0; q < nq; ++q)
                                if (r < n) RG(q, r) = b[r + (c0 + q) * ldb];
                            hi = (hi + 32 < n) ? hi + 32 : n;
                        }
                        __syncwarp();
                    }
                    const i64 p = __shfl_sync(0xffffffffu, pv, s);
                    if (p != j && lane < nq) {
                        const double t0 = RG(lane, j);
                        RG(lane, j) = RG(lane, p);
                        RG(lane, p) = t0;
                    }
                    __syncwarp();
                    double t[NR];
#pragma unroll
                    for (int q = 0; q < NR; ++q) t[q] = -RG(q, j);
#pragma unroll
                    for (int k = 0; k < KPL; ++k) {
                        const int i = 1 + lane + 32 * k;
                        if (i <= kl && j + i < n) {
#pragma unroll
                            for (int q = 0; q < NR; ++q)
                                if (q < nq) RG(q, j + i) = fma(t[q], L[s][k], RG(q, j + i));
                        }
                    }
                    __syncwarp();
                }
            }
#pragma unroll
            for (int s = 0; s < SB; ++s)
#pragma unroll
                for (int k = 0; k < KPL; ++k) L[s][k] = Ln[s][k];
            pv = pvn;
        }
        // retire what is still in the ring: rows from the last retired multiple of 32 to n-1
        i64 done = ((n - 2) >= 0) ? ((n - 2) & ~(i64)31) : 0;  // rows < done were retired
        for (int q = 0; q < nq; ++q)
            for (i64 r = done + lane; r < n; r += 32) b[r + (c0 + q) * ldb] = RG(q, r);
        __syncwarp();
    }

    // ---------------- backward: U x = y (DTBSV upper, no-trans, non-unit; bandwidth kv) ----------------
    if (!skip_u) {
        // rows [lo, n) resident
        i64 lo = (n - (kv + 64) > 0) ? n - (kv + 64) : 0;
        for (int q = 0; q < nq; ++q)
            for (i64 r = lo + lane; r < n; r += 32) RG(q, r) = b[r + (c0 + q) * ldb];
        __syncwarp();
        constexpr int KPU = 2 * KPL + 1;  // callers guarantee 32*KPU >= kv
        double U[SB][KPU], Un[SB][KPU], dg[SB], dgn[SB];
        auto loadU = [&](i64 jtop, double (&dst)[SB][KPU], double (&d)[SB]) {  // steps jtop, jtop-1, ..., jtop-SB+1
#pragma unroll
            for (int s = 0; s < SB; ++s) {
                const i64 j = jtop - s;
#pragma unroll
                for (int k = 0; k < KPU; ++k) {
                    const int i = 1 + lane + 32 * k;  // distance above the diagonal
                    dst[s][k] = (j >= 0 && i <= kv && j - i >= 0) ? ab[(kv - i) + j * ldab] : 0.0;
                }
                d[s] = (j >= 0) ? ab[kv + j * ldab] : 1.0;
            }
        };
        loadU(n - 1, U, dg);
        for (i64 jtop = n - 1; jtop >= 0; jtop -= SB) {
            loadU(jtop - SB, Un, dgn);
#pragma unroll
            for (int s = 0; s < SB; ++s) {
                const i64 j = jtop - s;
                if (j >= 0) {
                    // window maintenance every 32 steps (counted from the bottom row)
                    const i64 k = n - 1 - j;
                    if ((k & 31) == 0) {
                        if (k >= 32)
                            for (int q = 0; q < nq; ++q) {
                                const i64 r = j + 1 + lane;  // rows j+1 .. j+32 are final
                                b[r + (c0 + q) * ldb] = RG(q, r);
                            }
                        __syncwarp();
                        if (lo > 0) {
                            const i64 nlo = (lo - 32 > 0) ? lo - 32 : 0;
                            for (int q = 0; q < nq; ++q) {
                                const i64 r = nlo + lane;
                                if (r < lo) RG(q, r) = b[r + (c0 + q) * ldb];
                            }
                            lo = nlo;
                        }
                        __syncwarp();
                    }
                    double t[NR];
#pragma unroll
                    for (int q = 0; q < NR; ++q) t[q] = RG(q, j) / dg[s];
                    __syncwarp();
                    if (lane < nq) {
                        // lane q keeps its own quotient (all lanes computed all NR quotients identically)
                        double mine = t[0];
#pragma unroll
                        for (int q = 1; q < NR; ++q) mine = (lane == q) ? t[q] : mine;
                        RG(lane, j) = mine;
                    }
#pragma unroll
                    for (int kk = 0; kk < KPU; ++kk) {
                        const int i = 1 + lane + 32 * kk;
                        if (i <= kv && j - i >= 0) {
#pragma unroll
                            for (int q = 0; q < NR; ++q)
                                if (q < nq) RG(q, j - i) = fma(-t[q], U[s][kk], RG(q, j - i));
                        }
                    }
                    __syncwarp();
                }
            }
#pragma unroll
            for (int s = 0; s < SB; ++s) {
#pragma unroll
                for (int kk = 0; kk < KPU; ++kk) U[s][kk] = Un[s][kk];
                dg[s] = dgn[s];
            }
        }
        // retire rows [0, first retired row)
        const i64 steps = n;                                   // steps executed
        const i64 lastk = ((steps - 1) & ~(i64)31);           // k of the last maintenance point
        const i64 top = n - 1 - lastk;                         // j at that point; rows > top were retired iff lastk>=32
        for (int q = 0; q < nq; ++q)
            for (i64 r = lane; r <= ((lastk >= 32) ? top : n - 1); r += 32) b[r + (c0 + q) * ldb] = RG(q, r);
    }
#undef RG
}

// 'T' / 'C' (real): solve A^T X = B.  U^T forward substitution then L^T backward sweep with the
// inverse row interchanges.  Dot-product form (reference DTBSV-T / DGEMV-T); one warp per RHS column,
// operands straight from global/L2.  Not on a benchmark path: kept simple.
int bmb_cluster_solve(bmb200_ctx *h, int mode, i64 n, i64 kl, i64 ku, i64 nrhs, const double *dAB, i64 ldab, double *dB, i64 ldb);  // gbtrs_cluster.cu
int bmb_gbtrs_t_fast(bmb200_ctx *h, i64 n, i64 kl, i64 ku, i64 nrhs, const double *dAB, i64 ldab, const i64 *d_ipiv, double *dB, i64 ldb, int *u_done,
                     int *l_done);  // pb.cu

__global__ void __launch_bounds__(128)
gbtrs_t_kernel(i64 n, int kl, int ku, i64 nrhs, const double *__restrict__ ab, i64 ldab,
               const i64 *__restrict__ ipiv, double *__restrict__ b, i64 ldb, int skip_u)
{
    const int lane = threadIdx.x & 31;
    const i64 c = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= nrhs) return;
    const int kv = kl + ku;
    double *x = b + c * ldb;
    for (i64 j = skip_u ? n : 0; j < n; ++j) {  // U^T y = b
        double part = 0.0;
        const i64 i0 = (j - kv > 0) ? j - kv : 0;
        for (i64 i = i0 + lane; i < j; i += 32) part = fma(ab[(kv + i - j) + j * ldab], x[i], part);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        __syncwarp();
        if (lane == 0) x[j] = (x[j] - part) / ab[kv + j * ldab];
        __syncwarp();
    }
    if (kl > 0)
        for (i64 j = n - 2; j >= 0; --j) {  // L^T x = y, then undo the interchanges
            const i64 lm = (kl < n - 1 - j) ? kl : (n - 1 - j);
            double part = 0.0;
            for (i64 i = 1 + lane; i <= lm; i += 32) part = fma(ab[(kv + i) + j * ldab], x[j + i], part);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            __syncwarp();
            if (lane == 0) {
                x[j] = x[j] - part;
                const i64 p = ipiv[j] - 1;
                if (p != j) { const double t = x[j]; x[j] = x[p]; x[p] = t; }
            }
            __syncwarp();
        }
}

// ------------------------------------------------------------------------------------------------
// Wide bands (kl > 128): one CTA of 1024 threads per right-hand side, threads span the band, the active window of
// b lives in a shared-memory ring, the L / U column of the next steps is prefetched into registers.
// ------------------------------------------------------------------------------------------------
#define GW_THREADS 1024
#define GW_PF 4

template <int KPL, int KPU>
__global__ void __launch_bounds__(GW_THREADS, 1)
gbtrs_wide_kernel(i64 n, int kl, int ku, const double *__restrict__ ab, i64 ldab, const i64 *__restrict__ ipiv,
                  double *__restrict__ b, i64 ldb, int ring, int skip_u)
{
    extern __shared__ double rg[];
    const int tid = threadIdx.x, M = ring - 1;
    const int kv = kl + ku;
    double *x = b + (i64)blockIdx.x * ldb;
#define RGW(row) rg[(int)(row) & M]
    if (kl > 0) {
        i64 hi = ((i64)kl + 2 * GW_THREADS < n) ? (i64)kl + 2 * GW_THREADS : n;  // rows [.., hi) resident
        for (i64 r = tid; r < hi; r += GW_THREADS) RGW(r) = x[r];
        __syncthreads();
        double L[GW_PF][KPL];
        auto loadL = [&](i64 j, double (&dst)[KPL]) {
#pragma unroll
            for (int k = 0; k < KPL; ++k) {
                const int i = 1 + tid + GW_THREADS * k;
                dst[k] = (j < n - 1 && i <= kl && j + i < n) ? ab[(kv + i) + j * ldab] : 0.0;
            }
        };
#pragma unroll
        for (int s = 0; s < GW_PF; ++s) loadL(s, L[s]);
        for (i64 j0 = 0; j0 < n - 1; j0 += GW_PF) {
#pragma unroll
            for (int s = 0; s < GW_PF; ++s) {
                const i64 j = j0 + s;
                if (j < n - 1) {
                    if ((j & (GW_THREADS - 1)) == 0) {  // retire 1024 finished rows, pull 1024 new ones
                        if (j >= GW_THREADS) x[j - GW_THREADS + tid] = RGW(j - GW_THREADS + tid);
                        __syncthreads();
                        if (hi < n) {
                            if (hi + tid < n) RGW(hi + tid) = x[hi + tid];
                            hi = (hi + GW_THREADS < n) ? hi + GW_THREADS : n;
                        }
                        __syncthreads();
                    }
                    const i64 p = ipiv[j] - 1;
                    if (tid == 0 && p != j) { const double t0 = RGW(j); RGW(j) = RGW(p); RGW(p) = t0; }
                    __syncthreads();
                    const double t = -RGW(j);
#pragma unroll
                    for (int k = 0; k < KPL; ++k) {
                        const int i = 1 + tid + GW_THREADS * k;
                        if (i <= kl && j + i < n) RGW(j + i) = fma(t, L[s][k], RGW(j + i));
                    }
                    loadL(j + GW_PF, L[s]);  // refill this slot for step j+PF
                    __syncthreads();
                }
            }
        }
        const i64 done = ((n - 2) >= 0) ? ((n - 2) & ~(i64)(GW_THREADS - 1)) : 0;
        for (i64 r = done + tid; r < n; r += GW_THREADS) x[r] = RGW(r);
        __syncthreads();
    }
    if (!skip_u) {
        i64 lo = (n - ((i64)kv + 2 * GW_THREADS) > 0) ? n - ((i64)kv + 2 * GW_THREADS) : 0;
        for (i64 r = lo + tid; r < n; r += GW_THREADS) RGW(r) = x[r];
        __syncthreads();
        double U[GW_PF][KPU], dg[GW_PF];
        auto loadU = [&](i64 j, double (&dst)[KPU], double &d) {
#pragma unroll
            for (int k = 0; k < KPU; ++k) {
                const int i = 1 + tid + GW_THREADS * k;
                dst[k] = (j >= 0 && i <= kv && j - i >= 0) ? ab[(kv - i) + j * ldab] : 0.0;
            }
            d = (j >= 0) ? ab[kv + j * ldab] : 1.0;
        };
#pragma unroll
        for (int s = 0; s < GW_PF; ++s) loadU(n - 1 - s, U[s], dg[s]);
        for (i64 jt = n - 1; jt >= 0; jt -= GW_PF) {
#pragma unroll
            for (int s = 0; s < GW_PF; ++s) {
                const i64 j = jt - s;
                if (j >= 0) {
                    const i64 k = n - 1 - j;
                    if ((k & (GW_THREADS - 1)) == 0) {
                        if (k >= GW_THREADS) x[j + 1 + tid] = RGW(j + 1 + tid);
                        __syncthreads();
                        if (lo > 0) {
                            const i64 nlo = (lo - GW_THREADS > 0) ? lo - GW_THREADS : 0;
                            if (nlo + tid < lo) RGW(nlo + tid) = x[nlo + tid];
                            lo = nlo;
                        }
                        __syncthreads();
                    }
                    const double q = RGW(j) / dg[s];
                    __syncthreads();
                    if (tid == 0) RGW(j) = q;
#pragma unroll
                    for (int kk = 0; kk < KPU; ++kk) {
                        const int i = 1 + tid + GW_THREADS * kk;
                        if (i <= kv && j - i >= 0) RGW(j - i) = fma(-q, U[s][kk], RGW(j - i));
                    }
                    loadU(j - GW_PF, U[s], dg[s]);
                    __syncthreads();
                }
            }
        }
        const i64 lastk = ((n - 1) & ~(i64)(GW_THREADS - 1));
        const i64 top = n - 1 - lastk;
        for (i64 r = tid; r <= ((lastk >= GW_THREADS) ? top : n - 1); r += GW_THREADS) x[r] = RGW(r);
    }
#undef RGW
}

template <int KPL, int KPU>
static int launch_wide(bmb200_ctx *h, i64 n, i64 kl, i64 ku, i64 nrhs, const double *dAB, i64 ldab, const i64 *d_ipiv,
                       double *dB, i64 ldb, int skip_u)
{
    int ring = 4096;
    while (ring < kl + ku + 1 + 3 * GW_THREADS) ring <<= 1;
    const size_t smem = (size_t)ring * sizeof(double);
    BMB_CUDA(h, cudaFuncSetAttribute(gbtrs_wide_kernel<KPL, KPU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gbtrs_wide_kernel<KPL, KPU><<<(unsigned)nrhs, GW_THREADS, smem, h->stream>>>(n, (int)kl, (int)ku, dAB, ldab, d_ipiv,
                                                                                 dB, ldb, ring, skip_u);
    BMB_LAUNCH_CHECK(h);
    return 0;
}

template <int NR, int KPL, int SB>
static int launch_n(bmb200_ctx *h, i64 n, i64 kl, i64 ku, i64 nrhs, const double *dAB, i64 ldab, const i64 *d_ipiv,
                    double *dB, i64 ldb, int skip_u)
{
    const i64 kv = kl + ku;
    int ring = 128;
    while (ring < kv + 1 + 96) ring <<= 1;
    const size_t smem = (size_t)GBTRS_WARPS * NR * ring * sizeof(double);
    if (smem > 220 * 1024) {
        snprintf(h->err, sizeof(h->err), "dgbtrs: band (%lld,%lld) too wide for the shared-memory ring",
                 (long long)kl, (long long)ku);
        return BMB200_ERR_CUDA;
    }
    BMB_CUDA(h, cudaFuncSetAttribute(gbtrs_n_kernel<NR, KPL, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const i64 tiles = cdiv64(nrhs, NR);
    const i64 blocks = cdiv64(tiles, GBTRS_WARPS);
    gbtrs_n_kernel<NR, KPL, SB><<<(unsigned)blocks, GBTRS_WARPS * 32, smem, h->stream>>>(n, (int)kl, (int)ku, nrhs, dAB,
                                                                                        ldab, d_ipiv, dB, ldb, ring, skip_u);
    BMB_LAUNCH_CHECK(h);
    return 0;
}

extern "C" int bmb200_dgbtrs(bmb200_handle_t h, char trans, int64_t n, int64_t kl, int64_t ku, int64_t nrhs,
                             const double *dAB, int64_t ldab, const int64_t *d_ipiv, double *dB, int64_t ldb)
{
    if (!h) return -1;
    const bool tr = (trans == 'T' || trans == 't' || trans == 'C' || trans == 'c');
    if (!tr && !(trans == 'N' || trans == 'n')) return -2;
    if (n < 0) return -3;
    if (kl < 0) return -4;
    if (ku < 0) return -5;
    if (nrhs < 0) return -6;
    if (ldab < 2 * kl + ku + 1) return -8;
    if (ldb < imax64(1, n)) return -11;
    if (n == 0 || nrhs == 0) return 0;
    if (!dAB || !d_ipiv || !dB) return -7;
    DeviceGuard g(h->device);
    if (tr) {
        // the U^T half (and, for interchange-free factors, the L^T half) as column sweeps through the tuned 'N' machinery (pb.cu);
        // what is left -- the L^T half of a factorisation with interchanges -- runs the generic one-warp-per-RHS kernel below
        int u_done = 0, l_done = 0;
        const int rcf = bmb_gbtrs_t_fast(h, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb, &u_done, &l_done);
        if (rcf) return rcf;
        if (u_done && l_done) return 0;
        const i64 blocks = cdiv64(nrhs, 4);
        gbtrs_t_kernel<<<(unsigned)blocks, 128, 0, h->stream>>>(n, (int)kl, (int)ku, nrhs, dAB, ldab, d_ipiv, dB, ldb, u_done);
        BMB_LAUNCH_CHECK(h);
        return 0;
    }
    // narrow bands (kl <= 28, kl+ku <= 32; C1, C4): slot-scheduled sweeps, P steps per shuffle round (gbtrs_slot.cu)
    {
        const int rc = bmb_gbtrs_slot(h, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
        if (rc != 1) return rc;
    }
    // bands up to (63, 127-kl): window in registers across the lanes of a warp (gbtrs_shfl.cu)
    {
        const int rc = bmb_gbtrs_shfl(h, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
        if (rc != 1) return rc;
    }
    // beyond the register-window kernels: interchange-free factors (diagonally dominant / SPD systems) take the cluster
    // pipeline whatever the band width (measured at (100,100), n = 2^17, 4 RHS: 216 ms with the lane kernel below, which used
    // to take every band up to kl = 128)
    {
        const int rc = bmb_gbtrs_blocked(h, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
        if (rc != 1) return rc;
    }
    // KPL covers kl (forward) and 2*KPL+1 covers kv = kl+ku (backward): need 32*KPL >= kl and 32*(2KPL+1) >= kv
    i64 need = cdiv64(kl, 32);
    const i64 need_u = cdiv64(imax64(0, kl + ku - 32), 64);
    if (need_u > need) need = need_u;
    if (need < 1) need = 1;
    // Factors WITH interchanges beyond the register-window kernels.  Only the L sweep involves the pivots: the generic kernels
    // below run it alone (skip_u) and the U sweep -- a plain upper-triangular band solve, DTBSV('U','N','N') -- goes through the
    // cluster pipeline like an interchange-free solve (~69 instead of ~600-850 ns per column), bit-identical as before.
    const i64 kvv = kl + ku;
    const int split = (kl > 0 && n > 1 && !h->tune.gbtrs_nosplit) ? 1 : 0;
    int rcl;
    // few right-hand sides per warp when there are few in total, so that more SMs take part
    if (need <= 1) rcl = (nrhs >= 1024) ? launch_n<4, 1, 8>(h, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb, split)
                                        : launch_n<1, 1, 8>(h, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb, split);
    else if (need <= 2) rcl = launch_n<2, 2, 4>(h, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb, split);
    else if (need <= 4) rcl = launch_n<2, 4, 2>(h, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb, split);
    // wide bands with interchanges: one CTA per right-hand side
    else if (kl <= GW_THREADS && kl + ku <= 2 * GW_THREADS) rcl = launch_wide<1, 2>(h, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb, split);
    else if (kl <= 2 * GW_THREADS && kl + ku <= 4 * GW_THREADS) rcl = launch_wide<2, 4>(h, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb, split);
    else rcl = 1;
    if (rcl == 0 && split) {
        const int rcu = bmb_cluster_solve(h, 0, n, 0, kvv, nrhs, dAB, ldab, dB, ldb);  // U x = y
        if (rcu == 1) {  // the pipeline does not take this shape: the U sweep of the generic kernels (kl = 0: their L sweep is empty)
            if (need <= 4) return launch_n<2, 4, 2>(h, n, 0, kvv, nrhs, dAB, ldab, d_ipiv, dB, ldb, 0);
            return launch_wide<2, 4>(h, n, 0, kvv, nrhs, dAB, ldab, d_ipiv, dB, ldb, 0);
        }
        return rcu;
    }
    if (rcl != 1) return rcl;
    snprintf(h->err, sizeof(h->err), "dgbtrs: band (%lld,%lld) wider than (2048, 4096-kl) is not supported", (long long)kl,
             (long long)ku);
    return BMB200_ERR_CUDA;
}
