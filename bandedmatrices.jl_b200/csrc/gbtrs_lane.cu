// gbtrs_lane.cu -- multi-RHS band solve, "one lane = one right-hand side" (narrow bands, many RHS: config C4).
//
// Same arithmetic contract as gbtrs.cu (DGBTRS 'N'): forward sweep = row interchange then
// b[j+k] = fma(-b[j], L[k,j], b[j+k]); backward sweep = true division by the diagonal then
// b[j-k] = fma(-x[j], U[j-k,j], b[j-k]).  Each sweep is a chain of n dependent steps per right-hand side, so the
// only parallelism is across right-hand sides: a warp owns 32 RHS columns, lane q keeps the (band+1)-row window of
// ITS column in registers (static rotation: the step loop is unrolled band+1 times so every window index is a
// compile-time constant), and every step costs one DFMA per band entry and nothing else on the dependency chain.
//   * L / U columns and pivots are warp-uniform: they are streamed GL_PF columns ahead into a shared-memory ring
//     with cp.async and read back as broadcast LDS.128.
//   * B is column-major (one RHS = one 8n-byte column), so per-lane accesses would touch 32 different lines per
//     instruction.  Instead the warp moves B in 32-row x 1-RHS chunks: at step s it prefetches one chunk of RHS
//     (s mod 32) with ONE coalesced 256-byte cp.async and stores one finished chunk with one coalesced 256-byte
//     STG; the chunks pass through a per-RHS shared-memory ring with odd pitch (conflict-free for both the
//     lane = row and the lane = RHS access pattern).
//   * the row interchange partner is warp-uniform, so it is a uniform branch tree to a static register swap.
#include "common.cuh"

#define GL_CR 64                 // coefficient ring slots (columns)
#define GL_PF 48                 // columns prefetched ahead == cp.async groups allowed in flight
#define GL_LA 160                // B rows prefetched ahead (multiple of 32; >= window + 31 + GL_PF)
#define GL_RB 256                // B ring rows per RHS (power of two; >= GL_LA + 96)
#define GL_BP (GL_RB + 1)        // odd pitch

__device__ __forceinline__ void gl_cp8(void *smem_dst, const void *gsrc)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void gl_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void gl_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// swap w[ph] with w[ph+d] for a warp-uniform d in [LO, HI]: binary tree of uniform branches, static swaps
template <int W, int LO, int HI>
__device__ __forceinline__ void swap_dyn(double (&w)[W], const int ph, const int d)
{
    if constexpr (LO == HI) {
        const double t = w[ph];
        w[ph] = w[ph + LO];
        w[ph + LO] = t;
    } else {
        constexpr int MID = (LO + HI) / 2;
        if (d <= MID) swap_dyn<W, LO, MID>(w, ph, d);
        else swap_dyn<W, MID + 1, HI>(w, ph, d);
    }
}

// KW: compile-time bound on the band reach (kl forward, kl+ku backward); EXACT: the run-time reach equals KW.
// The window is w[0 .. KW+GL_U): phase ph of a GL_U-step block works on w[ph .. ph+KW]; after GL_U steps the
// registers are shifted down by GL_U (KW moves per GL_U steps), so the unrolled body is GL_U phases, not KW+1.
#define GL_U 8
template <int KW, bool FWD, bool EXACT>
__global__ void __launch_bounds__(32, 1)
gbtrs_lane(i64 n, int kl, int ku, i64 nrhs, const double *__restrict__ ab, i64 ldab, const i64 *__restrict__ ipiv,
           double *__restrict__ b, i64 ldb)
{
    constexpr int CP = KW + 2;                        // coefficient slot pitch (even => 16-byte aligned slots)
    constexpr int NCP = ((FWD ? KW : KW + 1) + 31) / 32;  // cp.async per lane per coefficient column
    extern __shared__ __align__(16) double sm[];
    double *cring = sm;                                    // GL_CR x CP : [0] diagonal (backward), [k] reach k
    double *bring = cring + GL_CR * CP;                    // 32 x GL_BP : per-RHS row ring
    long long *pring = (long long *)(bring + 32 * GL_BP);  // GL_RB      : pivots (forward)
    const int lane = threadIdx.x;
    const i64 r0 = (i64)blockIdx.x * 32;
    const int nq = (int)((nrhs - r0 < 32) ? (nrhs - r0) : 32);
    const int kv = kl + ku;
    const int nb = FWD ? kl : kv;
    const int ne = FWD ? kl : kv + 1;  // entries per coefficient column
    // virtual step / row index v = 0..n-1; the matrix row (= column of AB) behind it:
    auto rowof = [&](i64 v) -> i64 { return FWD ? v : n - 1 - v; };
    const i64 cstep = FWD ? ldab : -ldab;
    // lane's source inside a coefficient column and its slot offset: forward entry e -> L[e+1], backward e -> U reach e
    const double *cbase = ab + (FWD ? kv + 1 + lane : kv - lane);
    double *cdst = cring + (FWD ? lane + 1 : lane);
    auto fetch_col = [&](i64 c, const double *colp) {  // colp = cbase + rowof(c)*ldab
        if (c < n) {
            double *dst = cdst + (int)(c & (GL_CR - 1)) * CP;
#pragma unroll
            for (int i = 0; i < NCP; ++i)
                if (lane + 32 * i < ne) gl_cp8(dst + 32 * i, FWD ? colp + 32 * i : colp - 32 * i);
        }
    };
    double *myring = bring + lane * GL_BP;

    // ---- prologue: rows [0, GL_LA), columns [0, GL_PF), pivots [0, GL_LA) ----
    for (int q = 0; q < nq; ++q)
        for (int c = 0; c < GL_LA / 32; ++c) {
            const i64 v = 32 * c + lane;
            if (v < n) gl_cp8(bring + q * GL_BP + (int)v, b + (r0 + q) * ldb + rowof(v));
        }
    for (int c = 0; c < GL_PF; ++c) fetch_col(c, cbase + rowof(c) * ldab);
    if (FWD)
        for (int c = 0; c < GL_LA / 32; ++c) {
            const i64 v = 32 * c + lane;
            if (v < n) gl_cp8(pring + (int)v, ipiv + v);
        }
    gl_commit();
    gl_wait<0>();
    __syncwarp();

    double w[KW + GL_U];
#pragma unroll
    for (int k = 0; k < KW + GL_U; ++k) w[k] = (k < KW && k < n) ? myring[k] : 0.0;

    // running pointers (advanced every step): coefficient column s+GL_PF, chunk to prefetch, chunk to store
    const double *cptr = cbase + rowof(GL_PF) * ldab;
    const i64 rstep = FWD ? 1 : -1;
    const double *pld = b + r0 * ldb + rowof(GL_LA + lane);  // (q = 0, row blk + GL_LA + lane)
    double *pst = b + r0 * ldb + rowof(lane) - 32 * rstep;   // (q = 0, row blk - 32 + lane)
    const i64 wrap = 32 * rstep - 32 * ldb;

    for (i64 sb = 0; sb < n; sb += GL_U) {
        const i64 blk = sb & ~(i64)31;
        const int q0 = (int)sb & 24;  // == sb & 31 (sb is a multiple of GL_U = 8)
        // ---- data movement for later steps: GL_U coefficient columns, GL_U row chunks, 32 pivots; one group ----
        {
            const bool ld_rows = blk + GL_LA + lane < n;  // the prefetched chunk row exists
            const int ldslot = (int)((blk + GL_LA + lane) & (GL_RB - 1));
#pragma unroll
            for (int ph = 0; ph < GL_U; ++ph) {
                fetch_col(sb + ph + GL_PF, cptr);
                cptr += cstep;
                if (ld_rows && q0 + ph < nq) gl_cp8(bring + (q0 + ph) * GL_BP + ldslot, pld);
                pld += ldb;
            }
            if (FWD && q0 == 0) {
                const i64 v = sb + GL_LA + lane;
                if (v < n) gl_cp8(pring + (int)(v & (GL_RB - 1)), ipiv + v);
            }
            gl_commit();
            gl_wait<GL_PF / GL_U>();
            __syncwarp();  // columns sb.. / their pivots / the previous steps' finished rows are visible to every lane
            const int stslot = (int)((blk - 32 + lane) & (GL_RB - 1));
#pragma unroll
            for (int ph = 0; ph < GL_U; ++ph) {
                if (blk >= 32 && q0 + ph < nq) *pst = bring[(q0 + ph) * GL_BP + stslot];
                pst += ldb;
            }
            if (q0 == 24) { pld += wrap; pst += wrap; }
        }
        // ---- GL_U elimination steps: shared memory is only READ here, so the loads can be scheduled early ----
        double fin[GL_U];
#pragma unroll
        for (int ph = 0; ph < GL_U; ++ph) {
            const i64 s = sb + ph;
            fin[ph] = 0.0;
            if (s < n) {
                {   // the row that enters the window at this step
                    const i64 vn = s + KW;
                    w[ph + KW] = (vn < n) ? myring[(int)(vn & (GL_RB - 1))] : 0.0;
                }
                const double *cc = cring + (int)(s & (GL_CR - 1)) * CP;
                double t;
                if (FWD) {
                    const int d = (int)(pring[(int)(s & (GL_RB - 1))] - 1 - s);
                    if (d > 0 && d <= KW) swap_dyn<KW + GL_U, 1, KW>(w, ph, d);
                    t = w[ph];
                } else {
                    t = w[ph] / cc[0];
                }
                fin[ph] = t;
                const double nt = -t;
                const double2 *c2 = reinterpret_cast<const double2 *>(cc);
#pragma unroll
                for (int k2 = 0; k2 <= KW / 2; ++k2) {
                    const double2 c = c2[k2];  // reaches 2*k2, 2*k2+1
                    if (2 * k2 >= 1 && 2 * k2 <= KW && (EXACT || 2 * k2 <= nb))
                        w[ph + 2 * k2] = fma(nt, c.x, w[ph + 2 * k2]);
                    if (2 * k2 + 1 <= KW && (EXACT || 2 * k2 + 1 <= nb))
                        w[ph + 2 * k2 + 1] = fma(nt, c.y, w[ph + 2 * k2 + 1]);
                }
            }
        }
        // ---- finished rows leave through the ring; shift the window ----
#pragma unroll
        for (int ph = 0; ph < GL_U; ++ph)
            if (sb + ph < n) myring[(int)((sb + ph) & (GL_RB - 1))] = fin[ph];
#pragma unroll
        for (int k = 0; k < KW; ++k) w[k] = w[k + GL_U];
    }
    // ---- flush the finished rows that the per-step stores have not reached ----
    __syncwarp();
    {
        i64 v0 = ((n - 1) & ~(i64)31) - 32;
        if (v0 < 0) v0 = 0;
        for (int q = 0; q < nq; ++q)
            for (i64 v = v0 + lane; v < n; v += 32) b[(r0 + q) * ldb + rowof(v)] = bring[q * GL_BP + (int)(v & (GL_RB - 1))];
    }
}

template <int KW, bool FWD>
static int launch_lane(bmb200_ctx *h, i64 n, i64 kl, i64 ku, i64 nrhs, const double *dAB, i64 ldab, const i64 *d_ipiv,
                       double *dB, i64 ldb)
{
    const size_t smem = ((size_t)GL_CR * (KW + 2) + 32 * GL_BP + GL_RB) * sizeof(double);
    const unsigned blocks = (unsigned)cdiv64(nrhs, 32);
    const bool exact = (FWD ? kl : kl + ku) == KW;
    if (exact) {
        BMB_CUDA(h, cudaFuncSetAttribute(gbtrs_lane<KW, FWD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gbtrs_lane<KW, FWD, true><<<blocks, 32, smem, h->stream>>>(n, (int)kl, (int)ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
    } else {
        BMB_CUDA(h, cudaFuncSetAttribute(gbtrs_lane<KW, FWD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gbtrs_lane<KW, FWD, false><<<blocks, 32, smem, h->stream>>>(n, (int)kl, (int)ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
    }
    BMB_LAUNCH_CHECK(h);
    return 0;
}

// Returns 1 when this file does not cover the shape (caller falls through to the ring kernels of gbtrs.cu).
int bmb_gbtrs_lane(bmb200_ctx *h, i64 n, i64 kl, i64 ku, i64 nrhs, const double *dAB, i64 ldab, const i64 *d_ipiv,
                   double *dB, i64 ldb)
{
    const i64 kv = kl + ku;
    if (nrhs < 16 || kl > 32 || kv > 64 || n < 2) return 1;
    int rc = 0;
    if (kl > 0) {
        if (kl <= 2) rc = launch_lane<2, true>(h, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
        else if (kl <= 4) rc = launch_lane<4, true>(h, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
        else if (kl <= 8) rc = launch_lane<8, true>(h, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
        else if (kl <= 16) rc = launch_lane<16, true>(h, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
        else rc = launch_lane<32, true>(h, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
        if (rc) return rc;
    }
    if (kv <= 2) rc = launch_lane<2, false>(h, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
    else if (kv <= 4) rc = launch_lane<4, false>(h, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
    else if (kv <= 8) rc = launch_lane<8, false>(h, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
    else if (kv <= 16) rc = launch_lane<16, false>(h, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
    else if (kv <= 32) rc = launch_lane<32, false>(h, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
    else rc = launch_lane<64, false>(h, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
    return rc;
}
