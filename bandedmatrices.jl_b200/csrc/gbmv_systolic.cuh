// gbmv_systolic.cuh -- the narrow-band streaming gbmv body, shared by the single-GPU kernel (gbmv.cu) and the
// row-sharded multi-GPU kernel (sharded.cu).  See gbmv.cu for the design notes.
#pragma once
#include "common.cuh"

template <int W, int LDV>
__device__ __forceinline__ void load_col(const double *__restrict__ a, i64 lda, i64 c, bool valid, double (&col)[W])
{
    if (LDV == 8) {
        double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (valid) {
            ld_stream_v4(a + c * 8, v[0], v[1], v[2], v[3]);
            if (W > 4) ld_stream_v4(a + c * 8 + 4, v[4], v[5], v[6], v[7]);
        }
#pragma unroll
        for (int r = 0; r < W; ++r) col[r] = v[r];
    } else if (LDV == 4) {
        double v[4] = {0, 0, 0, 0};
        if (valid) ld_stream_v4(a + c * 4, v[0], v[1], v[2], v[3]);
#pragma unroll
        for (int r = 0; r < W; ++r) col[r] = v[r];
    } else if (LDV == 2) {
        double v[2] = {0, 0};
        if (valid) ld_stream_v2(a + c * 2, v[0], v[1]);
#pragma unroll
        for (int r = 0; r < W; ++r) col[r] = v[r];
    } else {
        const double *p = a + c * lda;
#pragma unroll
        for (int r = 0; r < W; ++r) col[r] = valid ? ld_stream(p + r) : 0.0;
    }
}

// x access policy of the single-GPU kernel: one contiguous vector.
struct XPlain {
    const double *x;
    __device__ __forceinline__ void prepare(i64, int) const {}
    __device__ __forceinline__ double stream(i64 c) const { return ld_stream(x + c); }
    __device__ __forceinline__ double plain(i64 c) const { return x[c]; }
};

// Virtual columns 0 .. m+ku-1 are cut into sets of 32 and runs of `sets_per_run` sets; run q is processed by warp
// (q mod nwarps).  Lane c of a set owns band column c; the W partial sums of a row hop lane -> lane+1 once per
// diagonal (carry[] hands them from lane 31 of one set to lane 0 of the next).  Rows completed by a run: [cs-ku, ce-ku).
template <int W, int LDV, class XSrc>
__device__ __forceinline__ void gbmv_n_systolic_body(i64 m, i64 n, int kl, int ku, double alpha,
                                                     const double *__restrict__ a, i64 lda, const XSrc xs, double beta,
                                                     double *__restrict__ y, i64 total_sets, i64 sets_per_run,
                                                     i64 num_runs)
{
    const int lane = threadIdx.x & 31;
    const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const i64 nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    const bool bz = (beta == 0.0);
    const int src = (lane + 31) & 31;

    for (i64 run = warp; run < num_runs; run += nwarps) {
        const i64 set0 = run * sets_per_run;
        const i64 set1 = (set0 + sets_per_run < total_sets) ? set0 + sets_per_run : total_sets;
        const i64 cs = set0 * 32;

        // ---- prologue: partial chains of the W-1 rows that started before column cs ----
        double carry[W > 1 ? W - 1 : 1];
        {
            double part = 0.0;
            if (lane < W - 1) {
                const i64 i = cs + kl - 1 - lane;  // row entering lane 0 of the first set at step lane+1
                if (i >= 0 && i < m) {
                    part = bz ? 0.0 : __dmul_rn(beta, y[i]);
                    i64 c0 = i - kl;
                    if (c0 < 0) c0 = 0;
                    i64 c1 = cs < n ? cs : n;
                    for (i64 c = c0; c < c1; ++c)
                        part = fma(__dmul_rn(alpha, xs.plain(c)), a[(ku + i - c) + c * lda], part);
                }
            }
#pragma unroll
            for (int s = 0; s < W - 1; ++s) carry[s] = shfl_d(part, s);
        }

        // ---- software-pipelined main loop: loads of set k+1 are in flight while set k is reduced ----
        double col[W], ncol[W];
        double xv, nxv = 0.0, yin = 0.0, nyin = 0.0;
        {
            const i64 c = cs + lane;
            const bool v = c < n;
            xs.prepare(cs, lane);
            load_col<W, LDV>(a, lda, c, v, col);
            xv = v ? xs.stream(c) : 0.0;
            if (!bz) yin = (c + kl < m) ? y[c + kl] : 0.0;
        }
        for (i64 set = set0; set < set1; ++set) {
            const i64 c = set * 32 + lane;
            if (set + 1 < set1) {
                const i64 cn = c + 32;
                const bool v = cn < n;
                xs.prepare(cn - lane, lane);
                load_col<W, LDV>(a, lda, cn, v, ncol);
                nxv = v ? xs.stream(cn) : 0.0;
                if (!bz) nyin = (cn + kl < m) ? y[cn + kl] : 0.0;
            }
            const bool valid = c < n;
            const double t = __dmul_rn(alpha, xv);
            double acc = bz ? 0.0 : __dmul_rn(beta, yin);
            if (valid) acc = fma(t, col[W - 1], acc);
#pragma unroll
            for (int s = 1; s < W; ++s) {
                const double prev = carry[s - 1];
                carry[s - 1] = acc;
                const double in = shfl_d(lane == 31 ? prev : acc, src);
                acc = valid ? fma(t, col[W - 1 - s], in) : in;
            }
            const i64 i = c - ku;
            if (i >= 0 && i < m) st_stream(y + i, acc);
#pragma unroll
            for (int r = 0; r < W; ++r) col[r] = ncol[r];
            xv = nxv;
            yin = nyin;
        }
    }
}

// Launch geometry shared by both users: an integer number of ~96-set runs per warp.
struct SystolicPlan {
    i64 total_sets, sets_per_run, num_runs, blocks;
};
static inline SystolicPlan systolic_plan(i64 m, i64 ku, int sm_count, int blocks_per_sm, int threads, int spr_override = 0)
{
    SystolicPlan p;
    p.total_sets = cdiv64(m + ku, 32);
    if (spr_override > 0) {  // short runs, one per warp, more blocks than are resident: the hardware block scheduler balances the tail
        p.sets_per_run = spr_override;
        p.num_runs = cdiv64(p.total_sets, p.sets_per_run);
        p.blocks = cdiv64(p.num_runs, threads / 32);
        return p;
    }
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    p.blocks = (i64)sm_count * blocks_per_sm;
    const i64 nwarps = p.blocks * (threads / 32);
    i64 k = p.total_sets / (nwarps * 96);
    if (k < 1) k = 1;
    p.sets_per_run = cdiv64(p.total_sets, nwarps * k);
    if (p.sets_per_run < 4) p.sets_per_run = 4;
    p.num_runs = cdiv64(p.total_sets, p.sets_per_run);
    if (p.num_runs < nwarps) p.blocks = cdiv64(p.num_runs, threads / 32);
    return p;
}
