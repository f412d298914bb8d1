// gbtrs_shfl.cu -- band solve with the active window in REGISTERS spread over the lanes of one warp.
//
// Same arithmetic contract as gbtrs.cu (DGBTRS 'N'): forward = row interchange, then
// b[j+k] = fma(-b[j], L[k,j], b[j+k]); backward = true division by the diagonal, then b[j-k] = fma(-x[j], U[j-k,j], b[j-k]).
// A sweep is a chain of n dependent steps; what decides the run time is the latency of ONE step, so the step is
// reduced to: one warp shuffle (broadcast of the pivot-row value) + one DFMA per lane.
//   * lane l holds the window rows r = l (mod 32) (RPL of them when the band reach exceeds 31), NR right-hand
//     sides per warp as independent chains;
//   * the row interchange is two shuffles (the value of row ipiv[j] is broadcast as the multiplier, the old row j
//     is handed to the lane that owned row ipiv[j]) -- no branch, no shared memory;
//   * the lane whose row retires keeps the finished value in a register; after 32 steps every lane holds one
//     finished row and the warp writes them with ONE coalesced store per RHS; rows entering the window are
//     prefetched 64 rows ahead with one coalesced cp.async per RHS per 32 steps;
//   * L / U columns (and pivots) are streamed GS_PF columns ahead into a shared-memory ring with cp.async and read
//     back one entry per lane.
#include "common.cuh"
#include <type_traits>

#define GS_CR 64   // coefficient ring slots (columns)
#define GS_PF 48   // columns prefetched ahead
#define GS_U 8     // steps per cp.async group
#define GS_BR 128  // B ring rows per RHS
#define GS_LA 64   // rows prefetched beyond the window

__device__ __forceinline__ void gs_cp8(void *smem_dst, const void *gsrc)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void gs_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void gs_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NR, int RPL, bool FWD>
__global__ void __launch_bounds__(32)
gbtrs_shfl(i64 n, int kl, int ku, i64 nrhs, const double *__restrict__ ab, i64 ldab, const i64 *__restrict__ ipiv,
           double *__restrict__ b, i64 ldb)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int WR = 32 * RPL;  // window rows
    constexpr int CP = WR + 2;    // coefficient slot pitch: entry k = reach k (k = 0: diagonal, backward only)
    extern __shared__ __align__(16) double sm[];
    double *cring = sm;                                      // GS_CR x CP
    double *bring = cring + GS_CR * CP;                      // NR x GS_BR
    long long *pring = (long long *)(bring + NR * GS_BR);    // GS_BR pivots (forward)
    const int lane = threadIdx.x;
    const i64 c0 = (i64)blockIdx.x * NR;
    const int nq = (int)((nrhs - c0 < NR) ? (nrhs - c0) : NR);
    const int kv = kl + ku;
    const int reach = FWD ? kl : kv;
    const int ne = FWD ? kl : kv + 1;
    auto rowof = [&](i64 v) -> i64 { return FWD ? v : n - 1 - v; };  // virtual step/row index -> matrix row
    const i64 cstep = FWD ? ldab : -ldab;
    const double *cbase = ab + (FWD ? kv + 1 + lane : kv - lane);    // lane's entry inside a coefficient column
    double *cdst = cring + (FWD ? lane + 1 : lane);
    auto fetch_col = [&](i64 c, const double *colp) {
        if (c < n) {
            double *dst = cdst + (int)(c & (GS_CR - 1)) * CP;
#pragma unroll
            for (int i = 0; i < RPL; ++i)
                if (lane + 32 * i < ne) gs_cp8(dst + 32 * i, FWD ? colp + 32 * i : colp - 32 * i);
        }
    };
    auto fetch_rows = [&](i64 vbase) {  // rows [vbase, vbase+32) of every RHS of this warp, pivots alongside
        const i64 v = vbase + lane;
        if (v < n) {
#pragma unroll
            for (int q = 0; q < NR; ++q)
                if (q < nq) gs_cp8(bring + q * GS_BR + (int)(v & (GS_BR - 1)), b + (c0 + q) * ldb + rowof(v));
        }
    };
    auto fetch_piv = [&](i64 vbase) {
        const i64 v = vbase + lane;
        if (FWD && v < n) gs_cp8(pring + (int)(v & (GS_BR - 1)), ipiv + v);
    };

    // ---- prologue ----
    double v[RPL][NR], fin[NR];
#pragma unroll
    for (int i = 0; i < RPL; ++i)
#pragma unroll
        for (int q = 0; q < NR; ++q) {
            const i64 r = lane + 32 * i;
            v[i][q] = (q < nq && r < n) ? b[(c0 + q) * ldb + rowof(r)] : 0.0;
        }
#pragma unroll
    for (int q = 0; q < NR; ++q) fin[q] = 0.0;
    for (int c = 0; c < GS_LA / 32; ++c) fetch_rows(WR + 32 * c);
    for (int c = 0; c < GS_LA / 32; ++c) fetch_piv(32 * c);
    for (int c = 0; c < GS_PF; ++c) fetch_col(c, cbase + rowof(c) * ldab);
    gs_commit();
    gs_wait<0>();
    __syncwarp();

    const double *cptr = cbase + rowof(GS_PF) * ldab;
    double nxt[NR];  // the row this lane takes when its slot-0 row retires (read once per 32 steps)
#pragma unroll
    for (int q = 0; q < NR; ++q) nxt[q] = 0.0;
    for (i64 sb = 0; sb < n; sb += GS_U) {
        // ---- data movement for later steps: GS_U coefficient columns (+ 32 rows / pivots every 32 steps) ----
#pragma unroll
        for (int ph = 0; ph < GS_U; ++ph) {
            fetch_col(sb + ph + GS_PF, cptr);
            cptr += cstep;
        }
        if ((sb & 31) == 0) {
            fetch_rows(sb + WR + GS_LA);
            fetch_piv(sb + GS_LA);
        }
        gs_commit();
        gs_wait<GS_PF / GS_U>();
        __syncwarp();
        const int a0 = (int)sb & 31;  // lane whose slot 0 holds row sb (sb is a multiple of GS_U: no wrap inside a batch)
        if (a0 == 0) {
            const i64 vn = sb + WR + lane;  // enters the window when row sb + lane retires
#pragma unroll
            for (int q = 0; q < NR; ++q) nxt[q] = (vn < n && q < nq) ? bring[q * GS_BR + ((int)vn & (GS_BR - 1))] : 0.0;
        }
        // ---- everything the GS_U steps read from shared memory, loaded up front (off the dependency chain) ----
        double cf[GS_U][RPL], dg[GS_U];
        int dd[GS_U];
#pragma unroll
        for (int ph = 0; ph < GS_U; ++ph) {
            const int sl = ((int)sb + ph) & (GS_CR - 1);
            const double *cc = cring + sl * CP;
            const int k0 = (lane - a0 - ph) & 31;
#pragma unroll
            for (int i = 0; i < RPL; ++i) cf[ph][i] = cc[k0 + 32 * i];
            dg[ph] = FWD ? 1.0 : cc[0];
            dd[ph] = FWD ? (int)(pring[((int)sb + ph) & (GS_BR - 1)] - 1 - (sb + ph)) : 0;
        }
        const int nsteps = (n - sb < GS_U) ? (int)(n - sb) : GS_U;
        auto run = [&](auto full_batch) {  // full batches carry no per-step bound check (branch-free chain)
#pragma unroll
        for (int ph = 0; ph < GS_U; ++ph) {
            if (decltype(full_batch)::value || ph < nsteps) {
                const int a = a0 + ph;
                const int k0 = (lane - a) & 31;
                bool on[RPL];
#pragma unroll
                for (int i = 0; i < RPL; ++i) on[i] = (k0 + 32 * i >= 1) && (k0 + 32 * i <= reach);
                const int bl = FWD ? ((a + dd[ph]) & 31) : a;  // lane holding the pivot row s + d
                const int hi = FWD ? (dd[ph] >> 5) : 0;       // ... in this slot
                const bool mine = lane == a;
#pragma unroll
                for (int q = 0; q < NR; ++q) {
                    double t;
                    if (FWD) {
                        double src = v[0][q];
#pragma unroll
                        for (int i = 1; i < RPL; ++i) src = (hi == i) ? v[i][q] : src;
                        t = __shfl_sync(FULL, src, bl);                   // value of the pivot row = multiplier
                        const double va = __shfl_sync(FULL, v[0][q], a);  // old row s goes where the pivot row was
                        if (lane == bl) {
#pragma unroll
                            for (int i = 0; i < RPL; ++i) v[i][q] = (hi == i) ? va : v[i][q];
                        }
                    } else {
                        t = __shfl_sync(FULL, v[0][q], a) / dg[ph];
                    }
                    fin[q] = mine ? t : fin[q];
                    const double nt = -t;
#pragma unroll
                    for (int i = 0; i < RPL; ++i)
                        if (on[i]) v[i][q] = fma(nt, cf[ph][i], v[i][q]);
                    if (mine) {  // row s retires: this lane's slots move up, the last one takes row s + WR
#pragma unroll
                        for (int i = 0; i + 1 < RPL; ++i) v[i][q] = v[i + 1][q];
                        v[RPL - 1][q] = nxt[q];
                    }
                }
            }
        }
        };
        if (nsteps == GS_U) run(std::true_type{});
        else run(std::false_type{});
        if (a0 == 24 && nsteps == GS_U) {  // every lane now holds one finished row of the last 32: coalesced stores
            const i64 r = rowof(sb - 24 + lane);
#pragma unroll
            for (int q = 0; q < NR; ++q)
                if (q < nq) b[(c0 + q) * ldb + r] = fin[q];
        }
    }
    if ((n & 31) != 0 && lane < (int)(n & 31)) {  // finished rows of the last partial block
        const i64 r = rowof((n & ~(i64)31) + lane);
#pragma unroll
        for (int q = 0; q < NR; ++q)
            if (q < nq) b[(c0 + q) * ldb + r] = fin[q];
    }
}

template <int NR, int RPL, bool FWD>
static int launch_shfl(bmb200_ctx *h, i64 n, i64 kl, i64 ku, i64 nrhs, const double *dAB, i64 ldab, const i64 *d_ipiv,
                       double *dB, i64 ldb)
{
    const size_t smem = ((size_t)GS_CR * (32 * RPL + 2) + (size_t)NR * GS_BR + GS_BR) * sizeof(double);
    const unsigned blocks = (unsigned)cdiv64(nrhs, NR);
    BMB_CUDA(h, cudaFuncSetAttribute(gbtrs_shfl<NR, RPL, FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gbtrs_shfl<NR, RPL, FWD><<<blocks, 32, smem, h->stream>>>(n, (int)kl, (int)ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
    BMB_LAUNCH_CHECK(h);
    return 0;
}

template <int NR, bool FWD>
static int launch_shfl_rpl(bmb200_ctx *h, i64 reach, i64 n, i64 kl, i64 ku, i64 nrhs, const double *dAB, i64 ldab,
                           const i64 *d_ipiv, double *dB, i64 ldb)
{
    if (reach <= 31) return launch_shfl<NR, 1, FWD>(h, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
    if (reach <= 63) return launch_shfl<NR, 2, FWD>(h, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
    return launch_shfl<NR, 4, FWD>(h, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
}

// Returns 1 when this file does not cover the shape (caller falls through to the other kernels of gbtrs.cu).
int bmb_gbtrs_shfl(bmb200_ctx *h, i64 n, i64 kl, i64 ku, i64 nrhs, const double *dAB, i64 ldab, const i64 *d_ipiv,
                   double *dB, i64 ldb)
{
    const i64 kv = kl + ku;
    if (kl > 63 || kv > 127) return 1;
    // right-hand sides per warp: one warp per SM sub-partition is the fastest chain; more RHS per warp only when
    // there are more right-hand sides than sub-partitions
    const i64 slots = (i64)h->sm_count * 4;
    const int nr = (nrhs <= slots) ? 1 : (nrhs <= 2 * slots) ? 2 : 4;
    int rc = 0;
    if (kl > 0) {
        if (nr == 1) rc = launch_shfl_rpl<1, true>(h, kl, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
        else if (nr == 2) rc = launch_shfl_rpl<2, true>(h, kl, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
        else rc = launch_shfl_rpl<4, true>(h, kl, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
        if (rc) return rc;
    }
    if (nr == 1) rc = launch_shfl_rpl<1, false>(h, kv, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
    else if (nr == 2) rc = launch_shfl_rpl<2, false>(h, kv, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
    else rc = launch_shfl_rpl<4, false>(h, kv, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
    return rc;
}
