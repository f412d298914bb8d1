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
#include <stdlib.h>

#include "common.cuh"

#define GP_NB 16          // panel width
#define GP_RPT 4          // panel rows per chain thread
#define GP_RING 32        // published L panels kept in the ring
#define GP_MAXNT 288      // threads per CTA (register budget of the chain role: 65536 / 288 = 227)
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
    int PY, PX;           // shared-memory pitches (chain: next-panel staging; update: tile)
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
template <int NB, int RPT>
__device__ void gp_chain(const PipeArgs &A, double *smem)
{
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, wid = tid >> 5, NW = NT >> 5;
    const int kl = A.kl, ku = A.ku, kv = kl + ku, R = NB + kl, PY = A.PY;
    const i64 ldab = A.ldab, n = A.n, m = A.m;
    double *const ab = A.ab;
    constexpr int CG = 8;
    constexpr int GPB = NB / CG;  // column groups per panel-aligned block
    // ---- shared memory carve ----
    double *Ys = smem;                                   // NB columns x PY: next panel's columns, rows J .. J+R+NB-1
    double *rowbuf = Ys + (size_t)NB * PY;               // [2][32][NB] best row of every warp
    double *posbuf = rowbuf + 2 * 32 * NB;               // [2][NB] row at the pivot position
    double *cand_val = posbuf + 2 * NB;                  // [2][32]
    double *cand_rinv = cand_val + 64;                   // [2][32]
    double *L11s = cand_rinv + 64;                       // [NB][NB]
    int *cand_row = (int *)(L11s + NB * NB);             // [2][32]
    int *s_piv = cand_row + 64;                          // [NB]
    volatile int *s_flag = (volatile int *)(s_piv + NB); // [0] ok, [1] prefetch ready

    double X[RPT][NB];
    long long ju = 0, ju_prev = 0;
    int info = 0;
    // ---- panel 0 straight from AB ----
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
        const int r = tid + q * NT;
#pragma unroll
        for (int c = 0; c < NB; ++c) X[q][c] = (r < R && r <= c + kl) ? ab[(i64)c * ldab + kv + r - c] : 0.0;
    }
    if (tid == 0) { s_flag[0] = 1; s_flag[1] = 0; }
    __syncthreads();

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
        int pf_have = 0;          // thread 0: flags already seen satisfied
        bool poll_inflight = false;
        int pollv = 0;
        auto issue_prefetch = [&]() {
#pragma unroll 1
            for (int c = 0; c < NB; ++c) {
                const i64 col = J + NB + c;
                const double *src = ab + col * ldab + (kv - NB - c);  // + r  ->  AB(kv + (J+r) - col, col)
                double *dst = Ys + (size_t)c * PY;
                for (int r = tid; r < R + NB; r += NT) {
                    const bool ok = (r <= NB + c + kl) && (J + r < m);
                    gp_cp_async8(dst + r, ok ? src + r : ab, ok);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        if (nwait == 0) {
            issue_prefetch();
            pf_issued = true;
        }

        // ================= B: factor the panel held in registers =================
#pragma unroll
        for (int jj = 0; jj < NB; ++jj) {
            const i64 j = J + jj;
            const int pb = jj & 1;
            // ---- thread-local candidate: first maximum over this thread's active rows ----
            double bsv = 0.0, bav = -1.0;
            int br = INT_MAX, bq = 0;
#pragma unroll
            for (int q = 0; q < RPT; ++q) {
                const int r = tid + q * NT;
                const double v = X[q][jj], av = fabs(v);
                if (r >= jj && r <= jj + kl && av > bav) { bav = av; bsv = v; br = r; bq = q; }
            }
            const double rown = 1.0 / bsv;  // own reciprocal, overlapped with the reductions
            bool win;
            const unsigned wrow = gp_argmax(br != INT_MAX ? (unsigned long long)__double_as_longlong(bav) : 0ull, (unsigned)br, win);
            if (win) {
                cand_val[pb * 32 + wid] = bsv;
                cand_rinv[pb * 32 + wid] = rown;
                cand_row[pb * 32 + wid] = (int)wrow;
                double *rb = rowbuf + (size_t)(pb * 32 + wid) * NB;
#pragma unroll
                for (int q = 0; q < RPT; ++q)
                    if (bq == q) {
#pragma unroll
                        for (int c = 0; c < NB; ++c) rb[c] = X[q][c];
                    }
            }
            if (tid == jj) {
#pragma unroll
                for (int c = 0; c < NB; ++c) posbuf[pb * NB + c] = X[0][c];
            }
            // ---- asynchronous poll of the next panel's progress flags (thread 0), result broadcast by the barrier ----
            if (!pf_issued && tid == 0 && jj >= 1) {
                if (poll_inflight && pollv >= k) ++pf_have;
                poll_inflight = false;
                if (pf_have >= nwait) {
                    __threadfence();
                    s_flag[1] = 1;
                } else {
                    pollv = ld_relaxed(A.prog + g0 + pf_have);
                    poll_inflight = true;
                }
            }
            __syncthreads();
            if (!pf_issued && s_flag[1] != 0) {
                issue_prefetch();
                pf_issued = true;
            }
            // ---- final reduction over the warps' candidates (every warp redundantly) ----
            const double cv = (lane < NW) ? cand_val[pb * 32 + lane] : 0.0;
            const double cri = (lane < NW) ? cand_rinv[pb * 32 + lane] : 0.0;
            const int cr = (lane < NW) ? cand_row[pb * 32 + lane] : INT_MAX;
            bool win2;
            const int p = (int)gp_argmax(cr != INT_MAX ? (unsigned long long)__double_as_longlong(fabs(cv)) : 0ull, (unsigned)cr, win2);
            const int wl = __ffs(__ballot_sync(FULLMASK, win2)) - 1;
            const double pv = shfl_d(cv, wl), rinv = shfl_d(cri, wl);
            const double *urow = rowbuf + (size_t)(pb * 32 + wl) * NB;
            if (tid == 0) {
                A.ipiv[j] = J + p + 1;
                slot_piv[jj] = p;
                s_piv[jj] = p;
            }
            double *gcol = ab + j * ldab + (kv - jj);  // gcol[r] = AB(kv + r - jj, j): multiplier of the row now at J+r
            if (pv != 0.0) {
                long long cand = j + ku + (p - jj);
                if (cand > n - 1) cand = n - 1;
                if (cand > ju) ju = cand;
                if (p != jj) {  // full-row interchange inside the panel
#pragma unroll
                    for (int q = 0; q < RPT; ++q)
                        if (tid + q * NT == p) {
#pragma unroll
                            for (int c = 0; c < NB; ++c) X[q][c] = posbuf[pb * NB + c];
                        }
                    if (tid == jj) {
#pragma unroll
                        for (int c = 0; c < NB; ++c) X[0][c] = urow[c];
                    }
                }
                double u[NB];
#pragma unroll
                for (int c = jj + 1; c < NB; ++c) u[c] = urow[c];
#pragma unroll
                for (int q = 0; q < RPT; ++q) {
                    const int r = tid + q * NT;
                    if (r > jj && r <= jj + kl) {
                        const double l = __dmul_rn(X[q][jj], rinv);
                        X[q][jj] = l;
                        gcol[r] = l;
#pragma unroll
                        for (int c = jj + 1; c < NB; ++c) X[q][c] = fma(-u[c], l, X[q][c]);
                    }
                }
            } else {
                if (info == 0) info = (int)(j + 1);
#pragma unroll
                for (int q = 0; q < RPT; ++q) {
                    const int r = tid + q * NT;
                    if (r > jj && r <= jj + kl) gcol[r] = X[q][jj];  // all zero; DGBTF2 leaves the column alone
                }
            }
        }
        // ================= publish panel k =================
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
                const int r = tid + q * NT;
                if (r >= NB && r < R) {
                    double *dst = L21 + (size_t)(r - NB) * NB;
#pragma unroll
                    for (int fk = 0; fk < 4; ++fk) stg4(dst + 4 * fk, X[q][fk], X[q][4 + fk], X[q][8 + fk], X[q][12 + fk]);
                }
            }
        }
        if (tid == 0) *(long long *)slot = ju;
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release(&A.ctl->panel_done, k + 1);
        ju_prev = ju;

        // ================= D: apply panel k to the next panel's columns (prefetched into Ys) =================
        if (!pf_issued) {
            for (int i = 0; i < nwait; ++i)
                if (!gp_cta_wait(A.prog + g0 + i, k, &A.ctl->abort, s_flag)) return;
            issue_prefetch();
            pf_issued = true;
        }
        gp_cp_async_wait_all();
        __syncthreads();
        if (tid == 0) s_flag[1] = 0;
        // ---- row interchanges + rows of U (forward substitution with unit-lower L11): one warp per column ----
        for (int c = wid; c < NB; c += NW) {
            double *y = Ys + (size_t)c * PY;
            if (lane == 0) {
                for (int jj = 0; jj < NB; ++jj) {
                    const int p = s_piv[jj];
                    if (p != jj) { const double t = y[jj]; y[jj] = y[p]; y[p] = t; }
                }
            }
            __syncwarp();
            double xi = (lane < NB) ? y[lane] : 0.0;
#pragma unroll
            for (int jj = 0; jj < NB - 1; ++jj) {
                const double uu = shfl_d(xi, jj);
                if (lane > jj && lane < NB) xi = fma(-uu, L11s[lane * NB + jj], xi);
            }
            if (lane < NB) {
                y[lane] = xi;
                ab[(J + NB + c) * ldab + (kv - NB - c) + lane] = xi;  // U12 row J+lane of column J+NB+c
            }
        }
        __syncthreads();
        // ---- Schur update of the rows below, L row in registers, U column broadcast from shared memory ----
#pragma unroll 1
        for (int c = 0; c < NB; ++c) {
            double *y = Ys + (size_t)c * PY;
            double u[NB];
#pragma unroll
            for (int jj = 0; jj < NB; ++jj) u[jj] = y[jj];
#pragma unroll
            for (int q = 0; q < RPT; ++q) {
                const int r = tid + q * NT;
                if (r >= NB && r < R) {
                    double acc = y[r];
#pragma unroll
                    for (int jj = 0; jj < NB; ++jj) acc = fma(-u[jj], X[q][jj], acc);
                    y[r] = acc;
                }
            }
        }
        __syncthreads();
        // ---- the updated columns become the next panel (rows shift by NB) ----
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
            const int r = tid + q * NT;
#pragma unroll
            for (int c = 0; c < NB; ++c) X[q][c] = (r < R) ? Ys[(size_t)c * PY + NB + r] : 0.0;
        }
        __syncthreads();
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
        if (tid == 0) { A.ctl->ju = ju; A.ctl->info = info; }
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

    for (int k = 0; k < A.KP; ++k) {
        if (!gp_cta_wait(&A.ctl->panel_done, k + 1, &A.ctl->abort, s_ok)) return;
        const double *slot = A.ring + (size_t)(k % GP_RING) * A.slot_doubles;
        const long long ju_k = __ldcg((const long long *)slot);
        if (tid < NB) s_piv[tid] = __ldcg((const int *)(slot + 1) + tid);
        for (int t = tid; t < NB * NB; t += NT) L11s[t] = __ldcg(slot + GP_HDR + t);
        __syncthreads();
        const double *L21 = slot + GP_HDR + NB * NB;
        const i64 J = (i64)k * NB;
        const i64 jmax = (ju_k < n - 1) ? ju_k : n - 1;
        const i64 g_first = (i64)(k + 2) * NB / CG, g_last = jmax / CG;
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
            // ---- interchanges + rows of U: one warp per column ----
            for (int q = wid; q < nc; q += NW) {
                double *y = Xs + (size_t)q * PX;
                const int rmin = (int)(c0 + q - kv - J);
                if (lane == 0) {
                    for (int jj = 0; jj < NB; ++jj) {
                        const int p = s_piv[jj];
                        if (p != jj && jj >= rmin) { const double t = y[jj]; y[jj] = y[p]; y[p] = t; }
                    }
                }
                __syncwarp();
                double xi = (lane < NB) ? y[lane] : 0.0;
#pragma unroll
                for (int jj = 0; jj < NB - 1; ++jj) {
                    const double uu = shfl_d(xi, jj);
                    if (lane > jj && lane < NB) xi = fma(-uu, L11s[lane * NB + jj], xi);
                }
                if (lane < NB) y[lane] = xi;
            }
            __syncthreads();
            // ---- Schur update on the FP64 tensor cores: X[NB + 8t .., :] += L21[8t .., :] * (-U12) ----
            double bf[G8][4];
#pragma unroll
            for (int g8 = 0; g8 < G8; ++g8)
#pragma unroll
                for (int s = 0; s < 4; ++s) bf[g8][s] = -Xs[(size_t)(8 * g8 + fr) * PX + 4 * s + fk];
            for (int t = wid; t < KLP / 8; t += NW) {
                double a0, a1, a2, a3;
                ldcg4(L21 + ((size_t)(8 * t + fr) * NB + 4 * fk), a0, a1, a2, a3);
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
            }
            __syncthreads();
            // ---- store the tile back ----
            for (int q = 0; q < nc; ++q) {
                const i64 c = c0 + q;
                const int rmin = (int)(c - kv - J);
                double *dstg = ab + c * ldab + (kv - (c - J));
                const double *srcs = Xs + (size_t)q * PX;
                for (int r = tid; r < R; r += NT)
                    if (r >= rmin) dstg[r] = srcs[r];
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) st_release(A.prog + g, k + 1);
        }
        if (tid == 0) st_release(A.done + u, k + 1);
    }
}

__global__ void __launch_bounds__(GP_MAXNT, 1) gbtrf_pipe_kernel(const PipeArgs A)
{
    extern __shared__ double gp_smem[];
    if (blockIdx.x == 0) gp_chain<GP_NB, GP_RPT>(A, gp_smem);
    else gp_update<GP_NB, 8>(A, gp_smem);
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
    static const bool off = getenv("BMB200_GBTRF_NOPIPE") != nullptr;
    if (off) return 0;
    constexpr int NB = GP_NB, CG = 8;
    const i64 R = NB + kl;
    if (kl < 32 || ku < 2 * NB || R > (i64)GP_RPT * GP_MAXNT) return 0;
    i64 KP = imin64((m - kl - NB) / NB + 1, n / NB - 1);
    if (m - kl - NB < 0) KP = 0;
    if (const char *e = getenv("BMB200_PIPE_MAXPANELS")) KP = imin64(KP, atoll(e));
    if (KP < 4) return 0;
    const int NT = (int)(((R + GP_RPT - 1) / GP_RPT + 31) / 32 * 32);
    const int KLP = (int)((kl + 7) & ~7);
    const int PY = (int)((R + NB + 1) & ~1);  // even: 16-byte aligned columns
    const int PX = pitch_mod16(NB + KLP, 2);
    const size_t smem_chain = ((size_t)NB * PY + 2 * 32 * NB + 2 * NB + 64 + 64 + NB * NB) * sizeof(double) + (64 + NB + 8) * sizeof(int);
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
    const size_t ctl_bytes = (sizeof(PipeCtl) + (size_t)(ngroups + U) * sizeof(int) + 255) & ~(size_t)255;
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
    a.PY = PY; a.PX = PX;
    void *args[] = {(void *)&a};
    BMB_CUDA(h, cudaLaunchCooperativeKernel((const void *)gbtrf_pipe_kernel, dim3(grid), dim3(NT), args, smem, h->stream));
    h->launches++;
    // control block back: abort flag, ju, info
    PipeCtl host;
    BMB_CUDA(h, cudaMemcpyAsync(&host, a.ctl, sizeof(PipeCtl), cudaMemcpyDeviceToHost, h->stream));
    BMB_CUDA(h, cudaStreamSynchronize(h->stream));
    if (host.abort) {
        snprintf(h->err, sizeof(h->err), "dgbtrf: pipelined kernel aborted (a flag wait expired; panel_done = %d of %lld)",
                 host.panel_done, (long long)KP);
        return BMB200_ERR_CUDA;
    }
    // PanelState of gbtrf_blocked.cu: {int info; int pad; long long ju;} in h->d_info
    struct { int info; int pad; long long ju; } st = {host.info, 0, host.ju};
    BMB_CUDA(h, cudaMemcpyAsync(h->d_info, &st, sizeof(st), cudaMemcpyHostToDevice, h->stream));
    BMB_CUDA(h, cudaStreamSynchronize(h->stream));
    *Jdone = KP * NB;
    return 0;
}
