// gbmm_wide.cu -- banded x banded for WIDE bands on the FP64 tensor cores: C <- alpha*A*B + beta*C (the product columns of
// _gbmm!, src/banded/gbmm.jl:296-340) when whole staged A columns no longer fit in shared memory (gbmm.cu's tile / ring
// kernels stop at about (64,64) x (64,64); before this file such products fell back to the scalar sweep kernel).
//
// A CTA owns a 64 x 64 tile of C in DENSE coordinates (rows k0.., columns j0..; only tiles that meet C's band exist) and walks
// the inner index v in blocks of KB = 32 over [max(k0-Al, j0-Bu), min(k0+63+Au, j0+63+Bl)] -- K-blocking, so shared memory
// holds two stages of a 64 x 32 slab of A and a 32 x 64 slab of B whatever the band widths are.  In band storage a column of A
// is contiguous in k and a column of B is contiguous in v, so both slabs are staged with coalesced cp.async runs; entries
// outside a band or outside the matrix are zero-filled (never read: NaN in the unused corners of the band arrays is harmless).
// Warp w owns the eight 8 x 8 tiles of tile row w; accumulators start from beta*C (or 0), every DMMA.8x8x4 adds four terms in
// ascending v (tools/fp64_peaks.cu: equal to the sequential FMA chain), K-blocks are walked in ascending v, the B operand is
// t = alpha*B[v,j] rounded first: every C[k,j] sees the FMAs of the reference's per-column dgbmv_ sequence in the same order
// (zero-filled terms add +0), so the result is bit-identical to the other kernels and to the oracle.  8 x 8 tiles whose own
// v range misses a K-block skip it (warp-uniform test).
// Shared-memory pitches: A slab sa[vv*72 + r] (72 = 8 mod 16) and B slab sb[c*36 + vv] (36 = 4 mod 16): a fragment load (4
// values of v x 8 rows / columns) touches every bank pair exactly twice, the minimum for 32 doubles.
#include "common.cuh"

#define GW_T 64
#define GW_KB 32
#define GW_PA 72
#define GW_PB 36
#define GW_THREADS 256

__device__ __forceinline__ void gw_dmma884(double &d0, double &d1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void gw_cp8(double *dst, const double *src, bool valid)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}

__global__ void __launch_bounds__(GW_THREADS)
gbmm_bb_kblock(i64 n, i64 nu, i64 mprod, int Al, int Au, int Bl, int Bu, int Cl, int Cu, double alpha, const double *__restrict__ a, i64 lda,
               const double *__restrict__ b, i64 ldb, double beta, double *__restrict__ c, i64 ldc, int RT)
{
    extern __shared__ __align__(16) double gw_sm[];
    auto sa = [&](int s) { return gw_sm + s * (GW_KB * GW_PA); };                       // two stages of the A slab
    auto sb = [&](int s) { return gw_sm + 2 * GW_KB * GW_PA + s * (GW_T * GW_PB); };    // two stages of the B slab
    const i64 jt = blockIdx.x / RT;
    const int rt = (int)(blockIdx.x - jt * RT);
    const i64 j0 = jt * GW_T;
    // tile rows of this tile column: those that meet rows [max(0, j0-Cu), min(n-1, j0+63+Cl)]
    const i64 klo = imax64_d(0, j0 - Cu), khi = imin64_d(n - 1, imin64_d(mprod - 1, j0 + GW_T - 1) + Cl);
    const i64 k0 = (klo / GW_T + rt) * GW_T;
    if (k0 > khi) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
    // inner range of the whole tile
    const i64 v_lo = imax64_d(imax64_d(k0 - Al, j0 - Bu), 0), v_hi = imin64_d(imin64_d(k0 + GW_T - 1 + Au, j0 + GW_T - 1 + Bl), nu - 1);
    // accumulators: lane holds C[k = k0 + 8*warp + lr][j = j0 + 8*t + 2*lc + q]
    const i64 kk = k0 + 8 * warp + lr;
    double acc[8][2];
#pragma unroll
    for (int t = 0; t < 8; ++t)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const i64 j = j0 + 8 * t + 2 * lc + q;
            const bool in = kk < n && j < mprod && kk - j <= Cl && j - kk <= Cu;
            acc[t][q] = (in && beta != 0.0) ? __dmul_rn(beta, c[(Cu + kk - j) + j * ldc]) : 0.0;
        }
    if (v_lo <= v_hi) {
        const i64 vb0 = v_lo & ~(i64)3;  // DMMA steps of 4 start at a multiple of 4 (any fixed grid keeps ascending order)
        const int nkb = (int)((v_hi - vb0) / GW_KB + 1);
        auto stage = [&](int s, i64 v0) {
            // A slab: (vv, r) -> A[k0 + r, v0 + vv], r fastest (contiguous in band storage)
            for (int e = tid; e < GW_KB * GW_T; e += GW_THREADS) {
                const int r = e & (GW_T - 1), vv = e >> 6;
                const i64 k = k0 + r, v = v0 + vv;
                const bool ok = v >= 0 && v < nu && k < n && k - v <= Al && v - k <= Au;
                gw_cp8(sa(s) + vv * GW_PA + r, a + (ok ? (Au + k - v) + v * lda : 0), ok);
            }
            // B slab: (cc, vv) -> B[v0 + vv, j0 + cc], vv fastest
            for (int e = tid; e < GW_KB * GW_T; e += GW_THREADS) {
                const int vv = e & (GW_KB - 1), cc = e / GW_KB;
                const i64 j = j0 + cc, v = v0 + vv;
                const bool ok = v >= 0 && v < nu && j < mprod && v - j <= Bl && j - v <= Bu;
                gw_cp8(sb(s) + cc * GW_PB + vv, b + (ok ? (Bu + v - j) + j * ldb : 0), ok);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        stage(0, vb0);
        // per 8 x 8 tile: its own inner range (tiles off the band have an empty one)
        const i64 kr0 = k0 + 8 * warp;
        for (int kb = 0; kb < nkb; ++kb) {
            const i64 v0 = vb0 + (i64)kb * GW_KB;
            if (kb + 1 < nkb) stage((kb + 1) & 1, v0 + GW_KB);
            else asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncthreads();
            const double *pa = sa(kb & 1) + lc * GW_PA + 8 * warp + lr;
            const double *pb = sb(kb & 1) + lr * GW_PB + lc;
            unsigned need = 0;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const i64 jc0 = j0 + 8 * t;
                const i64 tlo = imax64_d(kr0 - Al, jc0 - Bu), thi = imin64_d(kr0 + 7 + Au, jc0 + 7 + Bl);
                if (tlo <= thi && tlo < v0 + GW_KB && thi >= v0) need |= 1u << t;
            }
            if (need) {
#pragma unroll
                for (int ks = 0; ks < GW_KB; ks += 4) {
                    const double av = pa[ks * GW_PA];
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        if (need & (1u << t)) {
                            const double bv = __dmul_rn(alpha, pb[8 * t * GW_PB + ks]);
                            gw_dmma884(acc[t][0], acc[t][1], av, bv);
                        }
                    }
                }
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int t = 0; t < 8; ++t)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const i64 j = j0 + 8 * t + 2 * lc + q;
            if (kk < n && j < mprod && kk - j <= Cl && j - kk <= Cu) c[(Cu + kk - j) + j * ldc] = acc[t][q];
        }
}

// host side: called by bmb200_dgbmm_bb (gbmm.cu) for the product columns [0, mprod) when the staged-column kernels do not fit
int bmb_gbmm_wide(bmb200_ctx *h, i64 n, i64 nu, i64 mprod, i64 Al, i64 Au, i64 Bl, i64 Bu, i64 Cl, i64 Cu, double alpha, const double *dA,
                  i64 lda, const double *dB, i64 ldb, double beta, double *dC, i64 ldc)
{
    const i64 ntc = cdiv64(mprod, GW_T);
    const i64 RT = (Cl + Cu + GW_T - 1) / GW_T + 2;  // tile rows a tile column can meet
    if (ntc * RT >= ((i64)1 << 31)) return 1;
    const size_t smem = (size_t)(2 * GW_KB * GW_PA + 2 * GW_T * GW_PB) * sizeof(double);
    BMB_CUDA(h, cudaFuncSetAttribute(gbmm_bb_kblock, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gbmm_bb_kblock<<<(unsigned)(ntc * RT), GW_THREADS, smem, h->stream>>>(n, nu, mprod, (int)Al, (int)Au, (int)Bl, (int)Bu, (int)Cl, (int)Cu, alpha,
                                                                          dA, lda, dB, ldb, beta, dC, ldc, (int)RT);
    BMB_LAUNCH_CHECK(h);
    return 0;
}
