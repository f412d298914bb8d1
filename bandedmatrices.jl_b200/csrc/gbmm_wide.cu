// gbmm_wide.cu -- banded x banded for WIDE bands on the FP64 tensor cores: C <- alpha*A*B + beta*C (the product columns of
// _gbmm!, src/banded/gbmm.jl:296-340) when whole staged A columns no longer fit in shared memory (gbmm.cu's tile / ring
// kernels stop at about (64,64) x (64,64); before this file such products fell back to the scalar sweep kernel).
//
// A CTA owns a 64 x 64 tile of C in DENSE coordinates (rows k0.., columns j0..; only tiles that meet C's band exist) and walks
// the inner index v in blocks of KB = 32 over [max(k0-Al, j0-Bu), min(k0+63+Au, j0+63+Bl)] -- K-blocking, so shared memory
// holds two stages of a 64 x 32 slab of A and a 32 x 64 slab of B whatever the band widths are.  In band storage a column of A
// is contiguous in k and a column of B is contiguous in v, so both slabs are staged with coalesced cp.async runs; entries
// outside a band or outside the matrix are zero-filled (never read: NaN in the unused corners of the band arrays is harmless).
// Warp w owns a 16 x 32 block of the tile (2 x 4 tiles of 8 x 8: two A and four B fragments per k-step of 4); accumulators start from beta*C (or 0), every DMMA.8x8x4 adds four terms in
// ascending v (tools/fp64_peaks.cu: equal to the sequential FMA chain), K-blocks are walked in ascending v, the B operand is
// t = alpha*B[v,j] rounded first: every C[k,j] sees the FMAs of the reference's per-column dgbmv_ sequence in the same order
// (zero-filled terms add +0), so the result is bit-identical to the other kernels and to the oracle.  8 x 8 tiles whose own
// v range misses a K-block skip it (warp-uniform test).
// Shared-memory pitches: A slab sa[vv*68 + r] and B slab sb[c*36 + vv], both = 4 mod 16: a 64-bit fragment load is served per
// half-warp (4 values of v x 4 rows / columns), and pitch = 4 (mod 16) puts those 16 doubles into 16 different 8-byte banks
// (pitch 72 = 8 mod 16 measured 90 M bank conflicts on this kernel: the two v pairs of a half-warp fell on the same banks).
#include "common.cuh"

#define GW_T 64
#define GW_KB 32
#define GW_PA 68
#define GW_PB 36
#define GW_THREADS 256
#ifndef GW_NST
#define GW_NST 2   // cp.async stages
#endif
#ifndef GW_MINB
#define GW_MINB 2  // CTAs per SM the register allocation aims at
#endif

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

__global__ void __launch_bounds__(GW_THREADS, GW_MINB)
gbmm_bb_kblock(i64 n, i64 nu, i64 mprod, int Al, int Au, int Bl, int Bu, int Cl, int Cu, double alpha, const double *__restrict__ a, i64 lda,
               const double *__restrict__ b, i64 ldb, double beta, double *__restrict__ c, i64 ldc, int RT)
{
    extern __shared__ __align__(16) double gw_sm[];
    auto sa = [&](int s) { return gw_sm + s * (GW_KB * GW_PA); };                       // two stages of the A slab
    auto sb = [&](int s) { return gw_sm + GW_NST * GW_KB * GW_PA + s * (GW_T * GW_PB); };    // two stages of the B slab
    const i64 jt = blockIdx.x / RT;
    const int rt = (int)(blockIdx.x - jt * RT);
    const i64 j0 = jt * GW_T;
    // tile rows of this tile column: those that meet rows [max(0, j0-Cu), min(n-1, j0+63+Cl)]
    const i64 klo = imax64_d(0, j0 - Cu), khi = imin64_d(n - 1, imin64_d(mprod - 1, j0 + GW_T - 1) + Cl);
    const i64 k0 = (klo / GW_T + rt) * GW_T;
    if (k0 > khi) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
    const int wr = warp >> 1, wc = warp & 1;  // warp tile: 8 x 8 tiles (2*wr + {0,1}) x (4*wc + {0..3})
    // inner range of the whole tile
    const i64 v_lo = imax64_d(imax64_d(k0 - Al, j0 - Bu), 0), v_hi = imin64_d(imin64_d(k0 + GW_T - 1 + Au, j0 + GW_T - 1 + Bl), nu - 1);
    // accumulators: lane holds C[k = k0 + 16*wr + 8*a + lr][j = j0 + 32*wc + 8*t + 2*lc + q]
    double acc[2][4][2];
#pragma unroll
    for (int ar = 0; ar < 2; ++ar)
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const i64 kk = k0 + 16 * wr + 8 * ar + lr, j = j0 + 32 * wc + 8 * t + 2 * lc + q;
                const bool in = kk < n && j < mprod && kk - j <= Cl && j - kk <= Cu;
                acc[ar][t][q] = (in && beta != 0.0) ? __dmul_rn(beta, c[(Cu + kk - j) + j * ldc]) : 0.0;
            }
    if (v_lo <= v_hi) {
        const i64 vb0 = v_lo & ~(i64)3;  // DMMA steps of 4 start at a multiple of 4 (any fixed grid keeps ascending order)
        const int nkb = (int)((v_hi - vb0) / GW_KB + 1);
        // staging: thread -> fixed (r, vv0) of the A slab and (vv, cc0) of the B slab; in band storage A[k,v] = a[Au + k + v*(lda-1)]
        // and B[v,j] = b[Bu + v + j*(ldb-1)], so consecutive elements of a thread are a constant pointer step apart.  K-blocks
        // that lie inside both bands and the matrix (all but the edges of a wide band) copy without any predicate.
        const int ar_ = tid & (GW_T - 1), avv0 = tid >> 6;           // A: vv = avv0 + 4*i, i < KB/4
        const int bvv = tid & (GW_KB - 1), bcc0 = tid / GW_KB;       // B: cc = bcc0 + (THREADS/KB)*i, i < T*KB/THREADS
        constexpr int A_IT = GW_KB * GW_T / GW_THREADS, A_VSTEP = GW_THREADS / GW_T;
        constexpr int B_IT = GW_KB * GW_T / GW_THREADS, B_CSTEP = GW_THREADS / GW_KB;
        const i64 ak = k0 + ar_;
        const double *abase = a + Au + ak;                            // + v*(lda-1)
        const double *bbase = b + Bu + (j0 + bcc0) * (ldb - 1);       // + v + i*B_CSTEP*(ldb-1)
        auto stage = [&](int s, i64 v0) {
            double *da = sa(s) + avv0 * GW_PA + ar_, *db = sb(s) + bcc0 * GW_PB + bvv;
            const bool inside = v0 >= 0 && v0 + GW_KB - 1 < nu && k0 + GW_T - 1 < n && j0 + GW_T - 1 < mprod &&
                                k0 + GW_T - 1 - v0 <= Al && v0 + GW_KB - 1 - k0 <= Au && v0 + GW_KB - 1 - j0 <= Bl && j0 + GW_T - 1 - v0 <= Bu;
            const double *pa_ = abase + (v0 + avv0) * (lda - 1), *pb_ = bbase + (v0 + bvv);
            if (inside) {
#pragma unroll
                for (int i = 0; i < A_IT; ++i) gw_cp8(da + i * A_VSTEP * GW_PA, pa_ + (i64)i * A_VSTEP * (lda - 1), true);
#pragma unroll
                for (int i = 0; i < B_IT; ++i) gw_cp8(db + i * B_CSTEP * GW_PB, pb_ + (i64)i * B_CSTEP * (ldb - 1), true);
            } else {
#pragma unroll
                for (int i = 0; i < A_IT; ++i) {
                    const i64 v = v0 + avv0 + i * A_VSTEP;
                    const bool ok = v >= 0 && v < nu && ak < n && ak - v <= Al && v - ak <= Au;
                    gw_cp8(da + i * A_VSTEP * GW_PA, ok ? pa_ + (i64)i * A_VSTEP * (lda - 1) : a, ok);
                }
                const i64 v = v0 + bvv;
#pragma unroll
                for (int i = 0; i < B_IT; ++i) {
                    const i64 j = j0 + bcc0 + i * B_CSTEP;
                    const bool ok = v >= 0 && v < nu && j < mprod && v - j <= Bl && j - v <= Bu;
                    gw_cp8(db + i * B_CSTEP * GW_PB, ok ? pb_ + (i64)i * B_CSTEP * (ldb - 1) : b, ok);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
#pragma unroll
        for (int s0 = 0; s0 < GW_NST - 1; ++s0) {  // prologue: NST-1 stages in flight
            if (s0 < nkb) stage(s0, vb0 + (i64)s0 * GW_KB);
            else asm volatile("cp.async.commit_group;" ::: "memory");
        }
        const i64 kr0 = k0 + 16 * wr, jw0 = j0 + 32 * wc;
        for (int kb = 0; kb < nkb; ++kb) {
            const i64 v0 = vb0 + (i64)kb * GW_KB;
            if (kb + GW_NST - 1 < nkb) stage((kb + GW_NST - 1) % GW_NST, v0 + (i64)(GW_NST - 1) * GW_KB);
            else asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group %0;" ::"n"(GW_NST - 1) : "memory");
            __syncthreads();
            const double *pa = sa(kb % GW_NST) + lc * GW_PA + 16 * wr + lr;
            const double *pb = sb(kb % GW_NST) + (32 * wc + lr) * GW_PB + lc;
            // per 8 x 8 tile: its own inner range (tiles off the band have an empty one); bit 4*ar + t
            unsigned need = 0;
#pragma unroll
            for (int ar = 0; ar < 2; ++ar)
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const i64 r0 = kr0 + 8 * ar, jc0 = jw0 + 8 * t;
                    const i64 tlo = imax64_d(r0 - Al, jc0 - Bu), thi = imin64_d(r0 + 7 + Au, jc0 + 7 + Bl);
                    if (tlo <= thi && tlo < v0 + GW_KB && thi >= v0) need |= 1u << (4 * ar + t);
                }
            if (need == 0xffu) {  // interior: no predicates
#pragma unroll
                for (int ks = 0; ks < GW_KB; ks += 4) {
                    const double a0 = pa[ks * GW_PA], a1 = pa[ks * GW_PA + 8];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const double bv = __dmul_rn(alpha, pb[8 * t * GW_PB + ks]);
                        gw_dmma884(acc[0][t][0], acc[0][t][1], a0, bv);
                        gw_dmma884(acc[1][t][0], acc[1][t][1], a1, bv);
                    }
                }
            } else if (need) {
#pragma unroll
                for (int ks = 0; ks < GW_KB; ks += 4) {
                    const double a0 = pa[ks * GW_PA], a1 = pa[ks * GW_PA + 8];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const double bv = __dmul_rn(alpha, pb[8 * t * GW_PB + ks]);
                        if (need & (1u << t)) gw_dmma884(acc[0][t][0], acc[0][t][1], a0, bv);
                        if (need & (16u << t)) gw_dmma884(acc[1][t][0], acc[1][t][1], a1, bv);
                    }
                }
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int ar = 0; ar < 2; ++ar)
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const i64 kk = k0 + 16 * wr + 8 * ar + lr, j = j0 + 32 * wc + 8 * t + 2 * lc + q;
                if (kk < n && j < mprod && kk - j <= Cl && j - kk <= Cu) c[(Cu + kk - j) + j * ldc] = acc[ar][t][q];
            }
}

// host side: called by bmb200_dgbmm_bb (gbmm.cu) for the product columns [0, mprod) when the staged-column kernels do not fit
int bmb_gbmm_wide(bmb200_ctx *h, i64 n, i64 nu, i64 mprod, i64 Al, i64 Au, i64 Bl, i64 Bu, i64 Cl, i64 Cu, double alpha, const double *dA,
                  i64 lda, const double *dB, i64 ldb, double beta, double *dC, i64 ldc)
{
    const i64 ntc = cdiv64(mprod, GW_T);
    const i64 RT = (Cl + Cu + GW_T - 1) / GW_T + 2;  // tile rows a tile column can meet
    if (ntc * RT >= ((i64)1 << 31)) return 1;
    const size_t smem = (size_t)GW_NST * (GW_KB * GW_PA + GW_T * GW_PB) * sizeof(double);
    BMB_CUDA(h, cudaFuncSetAttribute(gbmm_bb_kblock, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gbmm_bb_kblock<<<(unsigned)(ntc * RT), GW_THREADS, smem, h->stream>>>(n, nu, mprod, (int)Al, (int)Au, (int)Bl, (int)Bu, (int)Cl, (int)Cu, alpha,
                                                                          dA, lda, dB, ldb, beta, dC, ldc, (int)RT);
    BMB_LAUNCH_CHECK(h);
    return 0;
}
