// gbtrf_pipe.cu -- wide-band partial-pivot LU as ONE persistent, flag-synchronised kernel (replaces LAPACK.gbtrf!,
// src/banded/BandedLU.jl:98, for bands too wide for the single-CTA window kernels).
//
// The factorisation is a chain of n dependent pivot steps, so the design goal is to keep that chain on one SM with
// nothing else on its critical path, and to hide all O(n kl ku) trailing-update work behind it on the other SMs:
//
//   CTA 0 ("chain")    owns the current NB-column panel in REGISTERS (thread = 4 panel rows x NB columns).  Per column:
//                      thread-local + REDUX pivot search, one barrier, pivot row broadcast through shared memory,
//                      multipliers (x * (1/pivot)) and the rank-1 update of the remaining panel columns as register
//                      FMAs.  Multipliers go to AB in LAPACK's un-permuted format as they are produced; the fully
//                      row-permuted L panel is published in a ring of slots, laid out as DMMA A-fragments.
//                      The chain CTA applies panel k to the NEXT panel's columns itself (from its registers, while the
//                      other CTAs are still busy with panel k-1), so it never waits for a round trip through another SM:
//                      the next panel's columns are prefetched with cp.async while the current panel is factored.
//   CTAs 1.. ("update") each owns the column groups g = u (mod U) of the trailing matrix.  For every published panel k
//                      and owned group inside (panel k+1, ju_k]: load the (NB+kl) x CG tile, apply the NB row
//                      interchanges, forward-substitute with the unit-lower L11 (rows of U), then the Schur update
//                      X -= L21 * U12 on the FP64 tensor cores (DMMA.8x8x4, A fragments streamed from the ring with one
//                      256-bit load per lane, C in shared memory), store, and publish per-group progress.
//
// Flags live in global memory (st.release / ld.relaxed + fence); every wait is a bounded spin that raises an abort flag
// instead of hanging.  All CTAs are co-resident (cooperative launch, one CTA per SM).
//
// Arithmetic contract (same as gbtrf.cu / gbtrf_blocked.cu): first-maximum pivots, reciprocal scaling, one FMA per
// element per eliminated column in ascending column order.  DMMA.8x8x4 accumulates k = 0..3 as a sequential FMA chain
// (verified bit-for-bit by tools/fp64_peaks.cu), and l*(-u) == (-u)*l exactly, so factors and pivots are bit-identical
// to DGBTF2 (LAPACK's blocked DGBTRF, which OpenBLAS runs for ku > 64, differs from that only by DGEMM rounding).
#include <climits>
#include <cooperative_groups.h>

#include <vector>

#include "common.cuh"

#define GP_NB 16          // panel width
#define GP_RPT 3          // panel rows per chain thread
#define GP_RING 32        // published L panels kept in the ring
#define GP_NT 384         // threads per CTA: 12 warps = 3 per scheduler; 65536 / 384 -> 168 registers per thread
#define GP_CG 8           // trailing columns per update work item
#define GP_HDR 16         // slot header, in doubles: [0] ju (i64), [1..8] NB relative pivot rows (int)
#define GP_SPIN_LIMIT (1u << 24)
#define FULLMASK 0xffffffffu

struct PipeCtl {          // device control block, zeroed before every launch
    int panel_done;       // panels published by the chain CTA
    int abort;            // a bounded spin expired somewhere
    int info;             // LAPACK info (first zero pivot, 1-based) after the pipelined panels
    int pad;
    long long ju;         // running 0-based ju after the pipelined panels
    long long pad2;
};

struct PipeArgs {
    i64 m, n;
    int kl, ku;
    double *ab;
    i64 ldab;
    i64 *ipiv;
    int KP;               // panels factored by this kernel
    PipeCtl *ctl;
    int *prog;            // per column group: panels applied so far (k+1 after panel k)
    int *done;            // per update CTA: panels completely applied
    double *ring;
    i64 slot_doubles;
    int KLP;              // kl rounded up to a multiple of 8
    int PY, PX;           // shared-memory pitches (chain: next-panel staging; update: tile)
    int speculate;        // 1: optimistic diagonal pivots with verification (falls back on the first violation)
    long long *stats;     // development counters (clock64 deltas): [0..15] chain, [16*(u+1) ..] update CTA u
};

// ---- small device helpers -----------------------------------------------------------------------------------------
__device__ __forceinline__ int ld_relaxed(const int *p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void ldcg4(const double *p, double &a, double &b, double &c, double &d)
{
    asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ void stg4(double *p, double a, double b, double c, double d)
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void gp_cp_async8(double *dst, const double *src, bool valid)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void gp_cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void gp_dmma(double &d0, double &d1, double a, double b);
// ---- TMA bulk copies (cp.async.bulk, UBLKCP) completing on an mbarrier: one instruction moves a whole column segment ----
__device__ __forceinline__ unsigned gp_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gp_mbar_init(unsigned long long *mbar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(count), "r"(gp_smem_u32(mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void gp_mbar_expect_tx(unsigned long long *mbar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(gp_smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ bool gp_mbar_try_wait(unsigned long long *mbar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                 : "=r"(ok)
                 : "r"(gp_smem_u32(mbar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void gp_mbar_wait(unsigned long long *mbar, unsigned parity)
{
    while (!gp_mbar_try_wait(mbar, parity)) {}
}
__device__ __forceinline__ void gp_bulk_g2s(void *sdst, const void *gsrc, unsigned bytes, unsigned long long *mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(gp_smem_u32(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(gp_smem_u32(mbar))
                 : "memory");
}
__device__ __forceinline__ void gp_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// A column segment of `len` doubles starting at gsrc (8-byte aligned) lands in the 16-byte aligned shared buffer `scol` so
// that element r is at scol[off + r], off = parity of gsrc's 8-byte slot: source and destination are then both 16-byte
// aligned after stepping one element back, and the size is rounded up to 16 bytes (one neighbouring element may be read
// on either side; callers keep that inside the allocation).  Returns the bytes the mbarrier has to expect.
__device__ __forceinline__ int gp_seg_off(const double *gsrc) { return (int)(((unsigned long long)gsrc >> 3) & 1ull); }
__device__ __forceinline__ unsigned gp_seg_bytes(const double *gsrc, int len) { return (unsigned)(((len + gp_seg_off(gsrc) + 1) & ~1) * 8); }
__device__ __forceinline__ void gp_seg_copy(double *scol, const double *gsrc, int len, unsigned long long *mbar)
{
    gp_bulk_g2s(scol, gsrc - gp_seg_off(gsrc), gp_seg_bytes(gsrc, len), mbar);
}

// Schur update of a shared-memory tile on the FP64 tensor cores: Y[c][NB + r] += sum_k L21[r][k] * (-Y[c][k]), r < KLP,
// for the 8*G8 columns at ycol(c) = ys + c*pitch + off[c].  L21 is read from a published ring slot (plane layout: plane
// fk holds (row, k = fk + 4s, s = 0..3) as 4 consecutive doubles), one 256-bit load per lane and row tile, with the next
// tile's fragments in flight.  DMMA.8x8x4 accumulates k ascending, so every element sees DGBTF2's FMA order.
template <int NB, int G8>
__device__ __forceinline__ void gp_schur_dmma(double *ys, int pitch, const int *off, const double *L21, int KLP, int wid, int lane, int NW)
{
    const int fr = lane >> 2, fk = lane & 3;
    double bf[G8][4];
    double *cp0[G8], *cp1[G8];
#pragma unroll
    for (int g8 = 0; g8 < G8; ++g8) {
        const int cb = 8 * g8 + fr;  // B fragment: (k = 4s + fk, column cb)
#pragma unroll
        for (int s = 0; s < 4; ++s) bf[g8][s] = -ys[(size_t)cb * pitch + off[cb] + 4 * s + fk];
        const int c0 = 8 * g8 + 2 * fk;  // C fragment: rows 8t + fr of columns c0, c0 + 1
        cp0[g8] = ys + (size_t)c0 * pitch + off[c0] + NB + fr;
        cp1[g8] = ys + (size_t)(c0 + 1) * pitch + off[c0 + 1] + NB + fr;
    }
    const int ntile = KLP / 8;
    const double *ap = L21 + ((size_t)fk * KLP + fr) * 4;  // plane fk, row fr; + 32 doubles per row tile
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0, n0 = 0, n1 = 0, n2 = 0, n3 = 0;
    int t = wid;
    if (t < ntile) ldcg4(ap + (size_t)t * 32, a0, a1, a2, a3);
    for (; t < ntile; t += NW) {
        if (t + NW < ntile) ldcg4(ap + (size_t)(t + NW) * 32, n0, n1, n2, n3);
#pragma unroll
        for (int g8 = 0; g8 < G8; ++g8) {
            double c0v = cp0[g8][8 * t], c1v = cp1[g8][8 * t];
            gp_dmma(c0v, c1v, a0, bf[g8][0]);
            gp_dmma(c0v, c1v, a1, bf[g8][1]);
            gp_dmma(c0v, c1v, a2, bf[g8][2]);
            gp_dmma(c0v, c1v, a3, bf[g8][3]);
            cp0[g8][8 * t] = c0v;
            cp1[g8][8 * t] = c1v;
        }
        a0 = n0; a1 = n1; a2 = n2; a3 = n3;
    }
}

__device__ __forceinline__ void gp_dmma(double &d0, double &d1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// thread 0 only: bounded spin until *flag >= target; acquires (fence + L1 invalidate) on success
__device__ bool gp_spin_ge(const int *flag, int target, int *abort_flag)
{
    unsigned it = 0;
    while (ld_relaxed(flag) < target) {
        if ((++it & 255u) == 0u && (ld_relaxed(abort_flag) != 0 || it > GP_SPIN_LIMIT)) {
            atomicExch(abort_flag, 1);
            return false;
        }
    }
    __threadfence();
    return true;
}
// whole CTA: wait for the flag; false => abort (uniform)
__device__ bool gp_cta_wait(const int *flag, int target, int *abort_flag, volatile int *s_ok)
{
    if (threadIdx.x == 0) *s_ok = gp_spin_ge(flag, target, abort_flag) ? 1 : 0;
    __syncthreads();
    const bool ok = *s_ok != 0;
    __syncthreads();
    return ok;
}

// warp-level "first maximum of |v|" over (key = bits of |v|, row): returns the winning row (INT_MAX if none) and whether
// this lane is the winner.  Inactive lanes pass key 0 / row INT_MAX.
__device__ __forceinline__ unsigned gp_argmax(unsigned long long key, unsigned row, bool &winner)
{
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mhi = __reduce_max_sync(FULLMASK, hi);
    const bool c1 = hi == mhi;
    const unsigned mlo = __reduce_max_sync(FULLMASK, c1 ? lo : 0u);
    const bool c2 = c1 && lo == mlo;
    const unsigned rmin = __reduce_min_sync(FULLMASK, c2 ? row : 0xffffffffu);
    winner = c2 && row == rmin;
    return rmin;
}

// ====================================================================================================================
// chain role
// ====================================================================================================================
template <int NB>
__device__ void gp_chain(const PipeArgs &A, double *smem)
{
    const bool want_spec = A.speculate != 0;
    constexpr int RPT = GP_RPT, NT = GP_NT, NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int kl = A.kl, ku = A.ku, kv = kl + ku, R = NB + kl, PY = A.PY;
    const i64 ldab = A.ldab, n = A.n, m = A.m;
    double *const ab = A.ab;
    constexpr int CG = GP_CG;
    constexpr int GPB = NB / CG;  // column groups per panel-aligned block
    // ---- shared memory carve ----
    double *Ys = smem;                                   // NB columns x PY: next panel's columns, rows J .. J+R+NB-1
    double *rowbuf = Ys + (size_t)NB * PY;               // [2][16][NB] best row of every warp (columns > jj)
    double *posbuf = rowbuf + 2 * 16 * NB;               // [2][NB] row at the pivot position
    double *cand_val = posbuf + 2 * NB;                  // [2][16]
    double *cand_rinv = cand_val + 32;                   // [2][16]
    double *L11s = cand_rinv + 32;                       // [NB][NB]
    double *swapbuf = L11s + NB * NB;                    // [NB] low columns of an interchanged pivot row
    int *cand_row = (int *)(swapbuf + NB);               // [2][16]
    int *s_piv = cand_row + 32;                          // [NB]
    volatile int *s_flag = (volatile int *)(s_piv + NB); // [0] ok, [1] prefetch ready
    int *s_off = s_piv + NB + 4;                         // [NB] 0/1 element offset of every staged column (16-byte alignment)
    unsigned long long *mbar = (unsigned long long *)(s_off + NB);  // completion barrier of the staged columns (8-byte aligned)
    unsigned pf_phase = 0;
    if (tid == 0) gp_mbar_init(mbar, 1);

    // ---- per-thread invariants: which of my rows take part in step jj (bit jj of the masks) ----
    unsigned actm[RPT], strm[RPT];   // candidate rows (jj <= r <= jj+kl) / eliminated rows (jj < r <= jj+kl)
    unsigned storem[RPT];            // bit c: (row r, panel column c) is a stored band entry (r < R, r <= c + kl)
    bool lowrow[RPT];                // rows NB <= r < R: Schur rows of the look-ahead update
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
        const int r = tid + q * NT;
        unsigned a = 0, s = 0;
        for (int jj = 0; jj < NB; ++jj) {
            if (r >= jj && r <= jj + kl) a |= 1u << jj;
            if (r > jj && r <= jj + kl) s |= 1u << jj;
        }
        actm[q] = a;
        strm[q] = s;
        unsigned w = 0;
        for (int c = 0; c < NB; ++c)
            if (r < R && r <= c + kl) w |= 1u << c;
        storem[q] = w;
        lowrow[q] = r >= NB && r < R;
    }

    double X[RPT][NB];
    long long ju = 0, ju_prev = 0;
    int info = 0;
    long long tPf = 0, tCk = 0, tSt = 0, tCm = 0, tB = 0, tPub = 0, tWait = 0, tCp = 0, tTrsm = 0, tSchur = 0, tReload = 0, tRing = 0, t0 = 0, nblock = 0;
#ifdef GP_PIPE_STATS   // development build: per-phase clock64 accounting by thread 0 (make NVFLAGS+=-DGP_PIPE_STATS)
#define GP_TICK(acc) do { if (tid == 0) { const long long t1_ = clock64(); acc += t1_ - t0; t0 = t1_; } } while (0)
#else
#define GP_TICK(acc) do { } while (0)
#endif
    if (tid == 0) t0 = clock64();
    // ---- panel 0 straight from AB ----
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
        const int r = tid + q * NT;
#pragma unroll
        for (int c = 0; c < NB; ++c) X[q][c] = (r < R && r <= c + kl) ? ab[(i64)c * ldab + kv + r - c] : 0.0;
    }
    if (tid == 0) { s_flag[0] = 1; s_flag[1] = 0; }
    __syncthreads();
    double *gbase = ab + kv + tid;   // AB(kv + r - jj, j) of row r = tid at the current step; advanced by ldab-1 per step
    bool speculate = want_spec;
    long long nopt = 0;
    int pollv0 = 0, pollv1 = 0;      // thread 0: progress flags of the next block, loaded one phase early

    for (int k = 0; k < A.KP; ++k) {
        const i64 J = (i64)k * NB;
        double *const slot = A.ring + (size_t)(k % GP_RING) * A.slot_doubles;
        int *const slot_piv = (int *)(slot + 1);
        // ---- ring back-pressure: slots k .. k+7 must have been consumed by every update CTA ----
        if (k >= GP_RING && (k & 7) == 0) {
            if (wid == 0) {
                const int need = k + 8 - GP_RING, U = (int)gridDim.x - 1;
                bool ok = true;
                unsigned it = 0;
                for (;;) {
                    int mn = INT_MAX;
                    for (int i = lane; i < U; i += 32) mn = min(mn, ld_relaxed(A.done + i));
                    mn = (int)__reduce_min_sync(FULLMASK, (unsigned)mn);
                    if (mn >= need) break;
                    if (++it > (GP_SPIN_LIMIT >> 4) || ld_relaxed(&A.ctl->abort) != 0) { ok = false; break; }
                }
                if (lane == 0) {
                    if (!ok) atomicExch(&A.ctl->abort, 1);
                    s_flag[0] = ok ? 1 : 0;
                }
            }
            __syncthreads();
            if (s_flag[0] == 0) return;
        }
        GP_TICK(tRing);
        // ---- prefetch bookkeeping: block k+1 must carry every update up to panel k-1 ----
        // (its groups exist as work items of panel k-1 only up to ju_{k-1}; columns beyond were never touched)
        const int g0 = (k + 1) * GPB;
        int nwait = 0;  // group flags to wait for
        if (k >= 1) {
            const i64 jmax = (ju_prev < n - 1) ? ju_prev : n - 1;
            for (int i = 0; i < GPB; ++i)
                if ((i64)(g0 + i) * CG <= jmax) nwait = i + 1;
        }
        bool pf_issued = false;
        auto issue_prefetch = [&]() {  // thread 0 only: NB bulk copies, rows J .. J+NB+c+kl of column J+NB+c
            gp_fence_proxy_async();    // Ys was last touched through the generic proxy (ordered by the preceding barrier)
            unsigned total = 0;
            const double *src = ab + (J + NB) * ldab + (kv - NB);  // AB(kv + J - col, col) for col = J+NB; + c*(ldab-1)
#pragma unroll 1
            for (int c = 0; c < NB; ++c) total += gp_seg_bytes(src + (i64)c * (ldab - 1), NB + c + kl + 1);
            gp_mbar_expect_tx(mbar, total);
#pragma unroll 1
            for (int c = 0; c < NB; ++c) {
                const double *sc = src + (i64)c * (ldab - 1);
                s_off[c] = gp_seg_off(sc);
                gp_seg_copy(Ys + (size_t)c * PY, sc, NB + c + kl + 1, mbar);
            }
        };
        // mode 0: use the flags loaded one phase ago; 1: poll now; 2: wait (bounded).  Uniform result, one barrier.
        auto prefetch_when_ready = [&](int mode) -> bool {
            if (tid == 0) {
                int ok = 1;
                if (mode == 0) {
                    if (nwait >= 1 && pollv0 < k) ok = 0;
                    if (nwait >= 2 && pollv1 < k) ok = 0;
                } else if (mode == 1) {
                    for (int i = 0; i < nwait; ++i)
                        if (ld_relaxed(A.prog + g0 + i) < k) ok = 0;
                } else {
                    for (int i = 0; i < nwait; ++i)
                        if (!gp_spin_ge(A.prog + g0 + i, k, &A.ctl->abort)) ok = -1;
                    if (ok > 0) ++nblock;
                }
                if (ok > 0) {
                    __threadfence();
                    issue_prefetch();
                }
                s_flag[1] = ok;
            }
            __syncthreads();
            const int v = s_flag[1];
            if (v > 0) pf_issued = true;
            return v >= 0;
        };
        if (!speculate) prefetch_when_ready(0);
        GP_TICK(tPf);
        bool anyswap = false;
        bool panel_done_opt = false;

        // ================= B (optimistic): assume every pivot is the diagonal, verify afterwards =================
        // For diagonally dominant systems (the 2-D Laplacian of the reference's example) partial pivoting never
        // interchanges.  The diagonal is LAPACK's pivot iff no row below it is strictly larger in magnitude (first
        // maximum), which every thread can check on its own rows: no reductions, no candidate broadcast.  The panel is
        // checkpointed in AB first; a single violation (or a zero pivot) restores it and falls back to the searching
        // path below for the rest of the factorisation, so pivots and factors are those of DGBTF2 in every case.
        if (speculate) {
            // No checkpoint is written: the staging buffer Ys still holds this panel as it was before step 0 (the next
            // block is only prefetched after the verification), so a violation just reloads the registers from it.
            int viol = 0;
            if (tid == 0) {  // publish row 0 for step 0
                const double d = X[0][0];
                cand_val[0] = d;
                cand_rinv[0] = 1.0 / d;
#pragma unroll
                for (int c = 1; c < NB; ++c) posbuf[c] = X[0][c];
            }
            auto ostep = [&](const int jj) {
                const int pb = jj & 1, nb = pb ^ 1;
                __syncthreads();
                const double pv = cand_val[pb * 16], rinv = cand_rinv[pb * 16];
                const double apv = fabs(pv);
                if (pv == 0.0) viol = 1;
                double l[RPT];
#pragma unroll
                for (int q = 0; q < RPT; ++q) {
                    const bool on = (strm[q] >> jj) & 1u;
                    const double x = X[q][jj];
                    if (on && fabs(x) > apv) viol = 1;
                    l[q] = on ? __dmul_rn(x, rinv) : 0.0;
                    if (on) X[q][jj] = l[q];
                }
                // column jj+1 first: the next pivot and its reciprocal (a ~130-cycle division) start while the
                // remaining columns are still being updated
                double dn = 1.0, rn = 1.0;
                if (jj + 1 < NB) {
                    const double uc = posbuf[pb * NB + jj + 1];
#pragma unroll
                    for (int q = 0; q < RPT; ++q) X[q][jj + 1] = fma(-uc, l[q], X[q][jj + 1]);
                    if (tid == jj + 1) {
                        dn = X[0][jj + 1];
                        rn = 1.0 / dn;
                    }
                }
#pragma unroll
                for (int c = jj + 2; c < NB; ++c) {
                    const double uc = posbuf[pb * NB + c];
#pragma unroll
                    for (int q = 0; q < RPT; ++q) X[q][c] = fma(-uc, l[q], X[q][c]);
                }
                if (jj + 1 < NB && tid == jj + 1) {
                    cand_val[nb * 16] = dn;
                    cand_rinv[nb * 16] = rn;
#pragma unroll
                    for (int c = jj + 2; c < NB; ++c) posbuf[nb * NB + c] = X[0][c];
                }
            };
#pragma unroll
            for (int jj = 0; jj < NB; ++jj) ostep(jj);
            if (tid == 0 && nwait > 0) {  // fresh progress flags for the prefetch decision below (latency hidden by the vote)
                pollv0 = ld_relaxed(A.prog + g0);
                pollv1 = ld_relaxed(A.prog + g0 + (GPB > 1 ? 1 : 0));
            }
            GP_TICK(tSt);
            if (__syncthreads_or(viol) == 0) {
                // commit: no interchanges, so LAPACK's un-permuted multipliers ARE the published panel; an update CTA
                // copies them from the ring slot to AB (flag in the slot header), keeping ~130 KB of narrow stores
                // per panel off this SM.  U11 goes to AB with the publication below.
                if (tid < NB) {
                    A.ipiv[J + tid] = J + tid + 1;
                    slot_piv[tid] = tid;
                    s_piv[tid] = tid;
                }
                if (tid == 0) slot_piv[NB] = 1;
                long long cand = J + (NB - 1) + ku;
                if (cand > n - 1) cand = n - 1;
                if (cand > ju) ju = cand;
                panel_done_opt = true;
                ++nopt;
                prefetch_when_ready(0);
            } else {
                speculate = false;  // restore the panel and search for pivots from here on
                if (k == 0) {
#pragma unroll
                    for (int q = 0; q < RPT; ++q) {
                        const int r = tid + q * NT;
#pragma unroll
                        for (int c = 0; c < NB; ++c) X[q][c] = (r < R && r <= c + kl) ? ab[(i64)c * ldab + kv + r - c] : 0.0;
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < NB; ++c) {
                        const double *yb = Ys + (size_t)c * PY + s_off[c] + NB + tid;
#pragma unroll
                        for (int q = 0; q < RPT; ++q) X[q][c] = (tid + q * NT < R) ? yb[q * NT] : 0.0;
                    }
                }
                __syncthreads();
                prefetch_when_ready(1);
            }
        }
        GP_TICK(tCm);
        if (!panel_done_opt) {
        gbase = ab + J * ldab + kv + tid;
        if (tid == 0) slot_piv[NB] = 0;

        // ================= B: factor the panel held in registers =================
        auto step = [&](const int jj) {
            const i64 j = J + jj;
            const int pb = jj & 1;
            // ---- thread-local candidate: first maximum over this thread's active rows ----
            // (compared through the bit pattern of |v|, like gp_argmax: a NaN ranks above every number, so a column
            // whose candidates are all NaN still yields a pivot row instead of "no candidate")
            double bsv = 0.0, bav = -1.0;
            long long bkey = -1;
            int bq = -1;
#pragma unroll
            for (int q = 0; q < RPT; ++q) {
                const double v = X[q][jj], av = fabs(v);
                const long long key = __double_as_longlong(av);
                if (((actm[q] >> jj) & 1u) && key > bkey) { bkey = key; bav = av; bsv = v; bq = q; }
            }
            const double rown = 1.0 / bsv;  // own reciprocal, overlapped with the reductions
            const int br = (bq >= 0) ? tid + bq * NT : INT_MAX;
            bool win;
            const unsigned wrow = gp_argmax(bq >= 0 ? (unsigned long long)__double_as_longlong(bav) : 0ull, (unsigned)br, win);
            if (win) {
                cand_val[pb * 16 + wid] = bsv;
                cand_rinv[pb * 16 + wid] = rown;
                cand_row[pb * 16 + wid] = (int)wrow;
                double *rb = rowbuf + (size_t)(pb * 16 + wid) * NB;
#pragma unroll
                for (int q = 0; q < RPT; ++q)
                    if (bq == q) {
#pragma unroll
                        for (int c = jj + 1; c < NB; ++c) rb[c] = X[q][c];
                    }
            }
            if (tid == jj) {
#pragma unroll
                for (int c = 0; c < NB; ++c) posbuf[pb * NB + c] = X[0][c];
            }
            __syncthreads();
            // ---- final reduction over the warps' candidates (every warp redundantly) ----
            const double cv = (lane < NW) ? cand_val[pb * 16 + lane] : 0.0;
            const double cri = (lane < NW) ? cand_rinv[pb * 16 + lane] : 0.0;
            const int cr = (lane < NW) ? cand_row[pb * 16 + lane] : INT_MAX;
            bool win2;
            const int p = (int)gp_argmax(cr != INT_MAX ? (unsigned long long)__double_as_longlong(fabs(cv)) : 0ull, (unsigned)cr, win2);
            const int wl = __ffs(__ballot_sync(FULLMASK, win2)) - 1;
            const double pv = shfl_d(cv, wl), rinv = shfl_d(cri, wl);
            const double *urow = rowbuf + (size_t)(pb * 16 + wl) * NB;
            if (tid == 0) {
                A.ipiv[j] = J + p + 1;
                slot_piv[jj] = p;
                s_piv[jj] = p;
            }
            if (pv != 0.0) {
                long long cand = j + ku + (p - jj);
                if (cand > n - 1) cand = n - 1;
                if (cand > ju) ju = cand;
                if (p != jj) {  // full-row interchange inside the panel (uniform branch)
                    anyswap = true;
#pragma unroll
                    for (int q = 0; q < RPT; ++q)
                        if (tid + q * NT == p) {
#pragma unroll
                            for (int c = 0; c < jj; ++c) swapbuf[c] = X[q][c];
#pragma unroll
                            for (int c = 0; c < NB; ++c) X[q][c] = posbuf[pb * NB + c];
                        }
                    __syncthreads();
                    if (tid == jj) {
#pragma unroll
                        for (int c = 0; c < jj; ++c) X[0][c] = swapbuf[c];
                        X[0][jj] = pv;
#pragma unroll
                        for (int c = jj + 1; c < NB; ++c) X[0][c] = urow[c];
                    }
                }
                // rows that take no part in this step (finished U rows, rows below the band, padding) get l = 0, so the
                // update below is an unconditional FMA: fma(-u, 0, x) == x (only the sign of a zero could change)
                double l[RPT];
#pragma unroll
                for (int q = 0; q < RPT; ++q) {
                    const bool on = (strm[q] >> jj) & 1u;
                    l[q] = on ? __dmul_rn(X[q][jj], rinv) : 0.0;
                    if (on) {
                        X[q][jj] = l[q];
                        gbase[q * NT] = l[q];
                    }
                }
#pragma unroll
                for (int c = jj + 1; c < NB; ++c) {
                    const double uc = urow[c];
#pragma unroll
                    for (int q = 0; q < RPT; ++q) X[q][c] = fma(-uc, l[q], X[q][c]);
                }
            } else {
                if (info == 0) info = (int)(j + 1);
#pragma unroll
                for (int q = 0; q < RPT; ++q)
                    if ((strm[q] >> jj) & 1u) gbase[q * NT] = X[q][jj];  // all zero; DGBTF2 leaves the column alone
            }
            gbase += ldab - 1;
        };
#pragma unroll
        for (int jj = 0; jj < NB / 2; ++jj) step(jj);
        if (!pf_issued) prefetch_when_ready(1);
#pragma unroll
        for (int jj = NB / 2; jj < NB; ++jj) step(jj);
        }
        GP_TICK(tB);
        // ================= publish panel k (the flag itself is released by warp 0 during the look-ahead) =================
        if (tid < NB) {
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                if (c >= tid) ab[(J + c) * ldab + kv + tid - c] = X[0][c];          // U11
                else { L11s[tid * NB + c] = X[0][c]; slot[GP_HDR + tid * NB + c] = X[0][c]; }  // unit-lower L11
            }
        }
        {
            double *L21 = slot + GP_HDR + NB * NB;
#pragma unroll
            for (int q = 0; q < RPT; ++q) {
                if (lowrow[q]) {  // plane fk holds (row, k = fk + 4s, s = 0..3): a warp writes 1 KB contiguous per plane
                    double *dst = L21 + (size_t)(tid + q * NT - NB) * 4;
#pragma unroll
                    for (int fk = 0; fk < 4; ++fk) stg4(dst + (size_t)fk * A.KLP * 4, X[q][fk], X[q][4 + fk], X[q][8 + fk], X[q][12 + fk]);
                }
            }
        }
        if (tid == 0) *(long long *)slot = ju;
        ju_prev = ju;
        GP_TICK(tPub);

        // ================= D: apply panel k to the next panel's columns (prefetched into Ys) =================
        if (!pf_issued && !prefetch_when_ready(2)) return;
        GP_TICK(tWait);
        gp_mbar_wait(mbar, pf_phase);   // the staged columns have landed (TMA writes are visible after the wait)
        pf_phase ^= 1u;
        if (tid < NB * NB) {  // rows below column c's band (r > NB+c+kl) are structural zeros; the copy left junk there
            const int c = tid >> 4, r = NB + c + kl + 1 + (tid & 15);
            if (r < R + NB + 1) Ys[(size_t)c * PY + s_off[c] + r] = 0.0;
        }
        __syncthreads();   // also orders every thread's slot / AB stores before the release below
        GP_TICK(tCp);
        if (wid == 0) {
            if (lane == 0) {
                st_release(&A.ctl->panel_done, k + 1);
                s_flag[1] = 0;
            }
        } else {
            // ---- row interchanges + rows of U (forward substitution with unit-lower L11): half a warp per column ----
            for (int c = 2 * (wid - 1) + (lane >> 4); c < NB; c += 2 * (NW - 1)) {
                double *y = Ys + (size_t)c * PY + s_off[c];
                const int i = lane & 15;
                if (anyswap) {
                    if (i == 0) {
                        for (int jj = 0; jj < NB; ++jj) {
                            const int p = s_piv[jj];
                            if (p != jj) { const double t = y[jj]; y[jj] = y[p]; y[p] = t; }
                        }
                    }
                    __syncwarp();
                }
                double xi = y[i];
                double lr[NB - 1];
#pragma unroll
                for (int jj = 0; jj < NB - 1; ++jj) lr[jj] = L11s[i * NB + jj];
#pragma unroll
                for (int jj = 0; jj < NB - 1; ++jj) {
                    const double uu = __shfl_sync(FULLMASK, xi, jj, 16);
                    if (i > jj) xi = fma(-uu, lr[jj], xi);
                }
                y[i] = xi;
                ab[(J + NB + c) * ldab + (kv - NB - c) + i] = xi;  // U12 row J+i of column J+NB+c
            }
        }
        __syncthreads();
        GP_TICK(tTrsm);
        // ---- Schur update of the rows below on the FP64 tensor cores, L21 read back from the slot just published ----
        gp_schur_dmma<NB, NB / 8>(Ys, PY, s_off, slot + GP_HDR + NB * NB, A.KLP, wid, lane, NW);
        __syncthreads();
        GP_TICK(tSchur);
        // ---- the updated columns become the next panel (rows shift by NB) ----
        {
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                const double *yb = Ys + (size_t)c * PY + s_off[c] + NB + tid;
#pragma unroll
                for (int q = 0; q < RPT; ++q) X[q][c] = (tid + q * NT < R) ? yb[q * NT] : 0.0;
            }
        }
        if (tid == 0 && k + 1 < A.KP) {  // progress flags of block k+2, consumed at the top of the next iteration
            pollv0 = ld_relaxed(A.prog + (k + 2) * GPB);
            pollv1 = ld_relaxed(A.prog + (k + 2) * GPB + (GPB > 1 ? 1 : 0));
        }
        __syncthreads();
        GP_TICK(tReload);
    }
    // ---- hand-over: the block after the last pipelined panel goes back to AB; ju / info for the tail kernels ----
    {
        const i64 J = (i64)A.KP * NB;
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
            const int r = tid + q * NT;
#pragma unroll
            for (int c = 0; c < NB; ++c)
                if (r < R && r <= c + kl && J + r < m) ab[(J + c) * ldab + kv + r - c] = X[q][c];
        }
        if (tid == 0) {
            A.ctl->ju = ju;
            A.ctl->info = info;
            long long *st = A.stats;
            st[0] = tB; st[1] = tPub; st[2] = tWait; st[3] = tCp; st[4] = tTrsm; st[5] = tSchur; st[6] = tReload; st[7] = tRing; st[8] = nblock; st[9] = nopt; st[10] = tPf; st[11] = tCk; st[12] = tSt; st[13] = tCm;
        }
    }
}

// ====================================================================================================================
// update role
// ====================================================================================================================
template <int NB, int CG>
__device__ void gp_update(const PipeArgs &A, double *smem)
{
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, wid = tid >> 5, NW = NT >> 5;
    const int kl = A.kl, ku = A.ku, kv = kl + ku, R = NB + kl, PX = A.PX, KLP = (kl + 7) & ~7;
    const i64 ldab = A.ldab, n = A.n;
    double *const ab = A.ab;
    const int u = (int)blockIdx.x - 1, U = (int)gridDim.x - 1;
    constexpr int G8 = CG / 8;
    double *Xs = smem;                                // CG columns x PX: rows J .. J+NB+KLP-1
    double *L11s = Xs + (size_t)CG * PX;              // [NB][NB]
    int *s_piv = (int *)(L11s + NB * NB);             // [NB]
    volatile int *s_ok = (volatile int *)(s_piv + NB);
    const int fr = lane >> 2, fk = lane & 3;
    long long tW = 0, tHdr = 0, tLoad = 0, tTr = 0, tMma = 0, tSt = 0, t0 = 0, nitems = 0;
    if (tid == 0) t0 = clock64();

    for (int k = 0; k < A.KP; ++k) {
        if (!gp_cta_wait(&A.ctl->panel_done, k + 1, &A.ctl->abort, s_ok)) return;
        GP_TICK(tW);
        const double *slot = A.ring + (size_t)(k % GP_RING) * A.slot_doubles;
        const long long ju_k = __ldcg((const long long *)slot);
        if (tid < NB) s_piv[tid] = __ldcg((const int *)(slot + 1) + tid);
        for (int t = tid; t < NB * NB; t += NT) L11s[t] = __ldcg(slot + GP_HDR + t);
        __syncthreads();
        bool anyswap = false;
#pragma unroll
        for (int jj = 0; jj < NB; ++jj) anyswap |= s_piv[jj] != jj;
        const int wb_flag = __ldcg((const int *)(slot + 1) + NB);
        const double *L21 = slot + GP_HDR + NB * NB;
        const i64 J = (i64)k * NB;
        const i64 jmax = (ju_k < n - 1) ? ju_k : n - 1;
        const i64 g_first = (i64)(k + 2) * NB / CG, g_last = jmax / CG;
        GP_TICK(tHdr);
        i64 g = g_first + (((i64)u - g_first) % U + U) % U;
        for (; g <= g_last; g += U) {
            const i64 c0 = g * CG;
            const int nc = (int)((jmax - c0 + 1 < CG) ? (jmax - c0 + 1) : CG);
            // ---- tile load: rows J .. J+R-1 of nc columns; slots above the stored band are structural zeros ----
            for (int q = 0; q < CG; ++q) {
                const i64 c = c0 + q;
                const int rmin = (int)(c - kv - J);
                const double *src = ab + c * ldab + (kv - (c - J));  // + r -> AB(kv + (J+r) - c, c)
                double *dst = Xs + (size_t)q * PX;
                for (int r = tid; r < R; r += NT) {
                    const bool ok = q < nc && r >= rmin;
                    gp_cp_async8(dst + r, ok ? src + r : ab, ok);
                }
            }
            gp_cp_async_wait_all();
            __syncthreads();
            GP_TICK(tLoad);
            // ---- interchanges + rows of U: half a warp per column ----
            for (int q = 2 * wid + (lane >> 4); q < CG; q += 2 * NW) {
                double *y = Xs + (size_t)q * PX;
                const int rmin = (int)(c0 + q - kv - J);
                const int i = lane & 15;
                if (anyswap) {
                    if (i == 0 && q < nc) {
                        for (int jj = 0; jj < NB; ++jj) {
                            const int p = s_piv[jj];
                            if (p != jj && jj >= rmin) { const double t = y[jj]; y[jj] = y[p]; y[p] = t; }
                        }
                    }
                    __syncwarp();
                }
                double xi = y[i];
                double lr[NB - 1];
#pragma unroll
                for (int jj = 0; jj < NB - 1; ++jj) lr[jj] = L11s[i * NB + jj];
#pragma unroll
                for (int jj = 0; jj < NB - 1; ++jj) {
                    const double uu = __shfl_sync(FULLMASK, xi, jj, 16);
                    if (i > jj) xi = fma(-uu, lr[jj], xi);
                }
                y[i] = xi;
            }
            __syncthreads();
            GP_TICK(tTr);
            // ---- Schur update on the FP64 tensor cores: X[NB + 8t .., :] += L21[8t .., :] * (-U12) ----
            double bf[G8][4];
#pragma unroll
            for (int g8 = 0; g8 < G8; ++g8)
#pragma unroll
                for (int s = 0; s < 4; ++s) bf[g8][s] = -Xs[(size_t)(8 * g8 + fr) * PX + 4 * s + fk];
            {
                const int ntile = KLP / 8;
                const double *ap = L21 + ((size_t)fk * KLP + fr) * 4;   // plane fk, row fr; + 32 doubles per row tile
                double a0 = 0, a1 = 0, a2 = 0, a3 = 0, n0 = 0, n1 = 0, n2 = 0, n3 = 0;
                int t = wid;
                if (t < ntile) ldcg4(ap + (size_t)t * 32, a0, a1, a2, a3);
                for (; t < ntile; t += NW) {
                    if (t + NW < ntile) ldcg4(ap + (size_t)(t + NW) * 32, n0, n1, n2, n3);  // next tile's A fragments in flight
#pragma unroll
                    for (int g8 = 0; g8 < G8; ++g8) {
                        double *cp = Xs + (size_t)(8 * g8 + 2 * fk) * PX + NB + 8 * t + fr;
                        double c0v = cp[0], c1v = cp[PX];
                        gp_dmma(c0v, c1v, a0, bf[g8][0]);
                        gp_dmma(c0v, c1v, a1, bf[g8][1]);
                        gp_dmma(c0v, c1v, a2, bf[g8][2]);
                        gp_dmma(c0v, c1v, a3, bf[g8][3]);
                        cp[0] = c0v;
                        cp[PX] = c1v;
                    }
                    a0 = n0; a1 = n1; a2 = n2; a3 = n3;
                }
            }
            __syncthreads();
            GP_TICK(tMma);
            // ---- store the tile back ----
            for (int q = 0; q < nc; ++q) {
                const i64 c = c0 + q;
                const int rmin = (int)(c - kv - J);
                double *dstg = ab + c * ldab + (kv - (c - J));
                const double *srcs = Xs + (size_t)q * PX;
                for (int r = tid; r < R; r += NT)
                    if (r >= rmin) dstg[r] = srcs[r];
            }
            __syncthreads();
            if (tid == 0) { st_release(A.prog + g, k + 1); ++nitems; }
            GP_TICK(tSt);
        }
        // ---- write-back of an interchange-free panel's multipliers from the ring slot to AB (LAPACK format) ----
        if (wb_flag != 0 && (k % U) == u) {
            const double *L21w = slot + GP_HDR + NB * NB;
            for (int r = tid; r < kl; r += NT) {  // panel row NB + r
                double v[NB];
#pragma unroll
                for (int fk2 = 0; fk2 < 4; ++fk2) ldcg4(L21w + ((size_t)fk2 * KLP + r) * 4, v[fk2], v[fk2 + 4], v[fk2 + 8], v[fk2 + 12]);
                double *dst = ab + J * ldab + kv + NB + r;  // AB(kv + (NB+r) - jj, J + jj) = dst[jj*(ldab-1)]
#pragma unroll
                for (int jj = 0; jj < NB; ++jj) {
                    if (NB + r <= jj + kl) *dst = v[jj];
                    dst += ldab - 1;
                }
            }
            for (int t = tid; t < NB * NB; t += NT) {
                const int i = t / NB, jj = t - i * NB;
                if (i > jj) ab[(J + jj) * ldab + kv + (i - jj)] = L11s[i * NB + jj];
            }
        }
        if (tid == 0) st_release(A.done + u, k + 1);
    }
    if (tid == 0) {
        long long *st = A.stats + 16 * (u + 1);
        st[0] = tW; st[1] = tHdr; st[2] = tLoad; st[3] = tTr; st[4] = tMma; st[5] = tSt; st[6] = nitems;
    }
}

__global__ void __launch_bounds__(GP_NT, 1) gbtrf_pipe_kernel(const PipeArgs A)
{
    extern __shared__ double gp_smem[];
    if (blockIdx.x == 0) gp_chain<GP_NB>(A, gp_smem);
    else gp_update<GP_NB, GP_CG>(A, gp_smem);
}

static int pitch_mod16(int need, int rem)  // smallest p >= need with p = rem (mod 16)
{
    int p = need;
    while ((p & 15) != rem) ++p;
    return p;
}

// Factors the first *Jdone columns (a multiple of NB) with the pipelined kernel and leaves AB, ipiv, and the PanelState
// in h->d_info (info, ju) exactly as the panel/update loop of gbtrf_blocked.cu would after the same panels.
// *Jdone = 0 when the shape is not eligible (the caller then runs everything with the stepwise kernels).
int bmb_gbtrf_pipe(bmb200_ctx *h, i64 m, i64 n, i64 kl, i64 ku, double *dAB, i64 ldab, i64 *d_ipiv, i64 *Jdone)
{
    *Jdone = 0;
    if (h->tune.gbtrf_nopipe) return 0;
    constexpr int NB = GP_NB, CG = GP_CG;
    const i64 R = NB + kl;
    if (kl < 32 || ku < 2 * NB || R > (i64)GP_RPT * GP_NT) return 0;
    // every pipelined panel k needs rows up to J + 2NB + kl - 1 and the whole block k+1 (+ one element of slack)
    i64 KP = imin64((m - kl - 2 * NB) / NB + 1, (n - 1) / NB - 1);
    if (m - kl - 2 * NB < 0) KP = 0;
    if (h->tune.pipe_maxpanels > 0) KP = imin64(KP, h->tune.pipe_maxpanels);
    if (KP < 4) return 0;
    const int NT = GP_NT;
    const int KLP = (int)((kl + 7) & ~7);
    const int PY = pitch_mod16((int)(R + NB + 3), 2);  // even (16-byte aligned columns, one spare element for the alignment shift); = 2 (mod 16): conflict-free C fragments
    const int PX = pitch_mod16(NB + KLP, 2);
    const size_t smem_chain = ((size_t)NB * PY + 2 * 16 * NB + 2 * NB + 32 + 32 + NB * NB + NB) * sizeof(double) + (32 + NB + 4 + NB + 4) * sizeof(int) + 16;
    const size_t smem_upd = ((size_t)CG * PX + NB * NB) * sizeof(double) + (NB + 8) * sizeof(int);
    const size_t smem = smem_chain > smem_upd ? smem_chain : smem_upd;
    if (smem > 225 * 1024) return 0;
    if (cudaFuncSetAttribute(gbtrf_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int per_sm = 0, coop = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gbtrf_pipe_kernel, NT, smem);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device);
    if (per_sm < 1 || !coop || h->sm_count < 8) return 0;
    const int grid = h->sm_count;  // one CTA per SM: chain + (SMs - 1) update CTAs, all co-resident
    const int U = grid - 1;
    const i64 ngroups = n / CG + 4;
    const i64 slot_doubles = GP_HDR + NB * NB + (i64)KLP * NB;
    const size_t flag_bytes = (sizeof(PipeCtl) + (size_t)(ngroups + U) * sizeof(int) + 255) & ~(size_t)255;
    const size_t stats_bytes = (size_t)16 * grid * sizeof(long long);
    const size_t ctl_bytes = (flag_bytes + stats_bytes + 255) & ~(size_t)255;
    const size_t ring_bytes = (size_t)GP_RING * slot_doubles * sizeof(double);
    int rc = bmb_ensure_scratch(h, ctl_bytes + ring_bytes);
    if (rc) return rc;
    BMB_CUDA(h, cudaMemsetAsync(h->scratch, 0, ctl_bytes + ring_bytes, h->stream));
    PipeArgs a;
    a.m = m; a.n = n; a.kl = (int)kl; a.ku = (int)ku; a.ab = dAB; a.ldab = ldab; a.ipiv = d_ipiv; a.KP = (int)KP;
    a.ctl = (PipeCtl *)h->scratch;
    a.prog = (int *)((char *)h->scratch + sizeof(PipeCtl));
    a.done = a.prog + ngroups;
    a.ring = (double *)((char *)h->scratch + ctl_bytes);
    a.slot_doubles = slot_doubles;
    a.PY = PY; a.PX = PX; a.KLP = KLP;
    a.speculate = h->tune.pipe_nospec ? 0 : 1;
    a.stats = (long long *)((char *)h->scratch + flag_bytes);
    void *args[] = {(void *)&a};
    const bool show = h->tune.pipe_stats != 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (show) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, h->stream); }
    BMB_CUDA(h, cudaLaunchCooperativeKernel((const void *)gbtrf_pipe_kernel, dim3(grid), dim3(NT), args, smem, h->stream));
    h->launches++;
    if (show) cudaEventRecord(e1, h->stream);
    // control block back: abort flag, ju, info
    PipeCtl host;
    BMB_CUDA(h, cudaMemcpyAsync(&host, a.ctl, sizeof(PipeCtl), cudaMemcpyDeviceToHost, h->stream));
    BMB_CUDA(h, cudaStreamSynchronize(h->stream));
    if (host.abort) {
        snprintf(h->err, sizeof(h->err), "dgbtrf: pipelined kernel aborted (a flag wait expired; panel_done = %d of %lld)",
                 host.panel_done, (long long)KP);
        return BMB200_ERR_CUDA;
    }
    if (show) {
        float kms = 0.f;
        cudaEventElapsedTime(&kms, e0, e1);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        fprintf(stderr, "[pipe] kernel %.3f ms for %lld panels (%.2f us/panel)\n", kms, (long long)KP, 1e3 * kms / (double)KP);
        std::vector<long long> st((size_t)16 * grid);
        cudaMemcpy(st.data(), a.stats, st.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        const double kp = (double)KP;
        fprintf(stderr, "[pipe] KP=%lld NT=%d smem=%zu | chain cycles/panel: B %.0f pub %.0f wait %.0f cp %.0f trsm %.0f schur %.0f reload %.0f ring %.0f | blocking waits %lld, optimistic panels %lld\n",
                (long long)KP, NT, smem, st[0] / kp, st[1] / kp, st[2] / kp, st[3] / kp, st[4] / kp, st[5] / kp, st[6] / kp, st[7] / kp, st[8], st[9]);
        fprintf(stderr, "[pipe] chain cycles/panel (before B): prefetch-check %.0f checkpoint %.0f opt-steps %.0f verify+commit %.0f\n", st[10] / kp, st[11] / kp, st[12] / kp, st[13] / kp);
        long long mx[7] = {0}, sm[7] = {0};
        for (int u = 1; u < grid; ++u)
            for (int i = 0; i < 7; ++i) { const long long v = st[(size_t)16 * u + i]; sm[i] += v; if (v > mx[i]) mx[i] = v; }
        const double it = sm[6] > 0 ? (double)sm[6] : 1.0;
        fprintf(stderr, "[pipe] update CTAs: items %lld (max/CTA %lld) | cycles/item: load %.0f trsm %.0f mma %.0f store %.0f | per CTA per panel: wait %.0f hdr %.0f\n",
                sm[6], mx[6], sm[2] / it, sm[3] / it, sm[4] / it, sm[5] / it, sm[0] / kp / U, sm[1] / kp / U);
    }
    // PanelState of gbtrf_blocked.cu: {int info; int pad; long long ju;} in h->d_info
    struct { int info; int pad; long long ju; } st = {host.info, 0, host.ju};
    BMB_CUDA(h, cudaMemcpyAsync(h->d_info, &st, sizeof(st), cudaMemcpyHostToDevice, h->stream));
    BMB_CUDA(h, cudaStreamSynchronize(h->stream));
    *Jdone = KP * NB;
    return 0;
}
