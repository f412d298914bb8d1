// common.cuh -- shared internals of libbmb200 (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/bmb200.h"
#include "../../include/bmb200_internal.h"

typedef int64_t i64;

struct bmb200_halo {
    // this rank's mailbox (device memory, exported through CUDA IPC):
    //   [0 .. max_halo)            x entries pushed by the LEFT neighbour  (its last kl entries)
    //   [max_halo .. 2*max_halo)   x entries pushed by the RIGHT neighbour (its first ku entries)
    //   flags: 2 x uint64 epoch counters at the end
    double *box = nullptr;
    double *left_box = nullptr;   // peer mapping of the left neighbour's mailbox
    double *right_box = nullptr;  // peer mapping of the right neighbour's mailbox
    i64 max_halo = 0;
    int rank = 0, nranks = 1;
    unsigned long long epoch = 0;
};

// Development knobs (A/B timing, diagnostics).  Release dispatch reads ONLY this block -- never the environment; the
// defaults below are the shipped behaviour and bmb200_internal_set_tuning (include/bmb200_internal.h) is the only writer.
struct bmb_tuning {
    int gbmm_ring = -1;        // 1 forces the persistent ring kernel for banded x banded, 0 disables it
    int gbmm_nt = 3;           // row tiles per work item (2 or 3)
    int gbmm_rw = 0;           // ring kernel warps per CTA (0 = by tile width)
    int gbmm_wide = -1;        // 1 forces the K-blocked wide-band kernel (gbmm_wide.cu), 0 disables it
    int gbtrf_nopipe = 0;      // 1 disables the pipelined wide-band LU
    int gbtrf_nomw = 0;        // 1 disables the multi-warp narrow-band LU (gbtrf_mw.cu)
    int gbtrf_nostrip = 0;     // 1 disables the strip-resident interchange-free LU
    long long pipe_maxpanels = 0;  // > 0 caps the panels the pipelined LU takes
    int pipe_nospec = 0;       // 1 disables optimistic diagonal pivoting
    int pipe_stats = 0;        // 1 prints the chain CTA's cycle breakdown
    int gbtrs_noblock = 0;     // 1 disables the panel-blocked interchange-free solve
    int gbtrs_pfdist_blocked = 1;
    int gbtrs_nosplit = 0;     // 1: factors with interchanges keep both sweeps on the generic kernels (no cluster U sweep)
    int gbtrs_nocluster = 0;   // 1 disables the cluster solve
    int gbtrs_cluster = 0;     // > 0 forces the cluster size
    int gbtrs_pfdist = 6;      // L2 prefetch distance (panels) of the cluster solve
    int gbtrs_stats = 0;
    int debug = 0;
    int typed_nowin = 0;       // 1 sends the S/C/Z band LU to the global-memory kernel even when the shared-memory window fits
    int gbmv_spr = 0;          // > 0: systolic gbmv in short runs of this many 32-column sets, one per warp, non-persistent grid
    int pb_nodiag = 0;         // -1: narrow-band dpbtrf (kd <= 31) takes the one-warp register kernel instead of the window kernel
    int pb_nopdl = 0;          // 1 captures the blocked Cholesky without programmatic dependent launch
    int pb_clate = 0;          // 1: the update kernel loads its C tile after the dependent-launch wait instead of before it
    int pb_nobulk = 0;         // 1: the blocked Cholesky's update kernel stages every slab by 8-byte cp.async (no bulk copies)
    int sbmv_rows_k = 16;      // band width below which dsbmv uses the row kernel
};

struct bmb200_ctx {
    bmb_tuning tune;
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    int64_t launches = 0;
    int last_gbmm_path = 0;    // product columns of the last bmb200_dgbmm_bb: 0 sweep, 1 tile DMMA, 2 ring DMMA, 3 K-blocked DMMA
    char err[256] = {0};
    // small persistent device scratch (LU bookkeeping, info words)
    int *d_info = nullptr;     // [0]=info, [1]=ju, spare
    void *scratch = nullptr;   // grow-only workspace
    size_t scratch_bytes = 0;
    void *backup = nullptr;    // grow-only copy of a band for the optimistic LU of an in-place call (gbtrf_strip.cu)
    size_t backup_bytes = 0;
    const double *lu_src = nullptr;  // set by bmb200_dgbtrf_from for the duration of the call: the un-widened source
    int64_t lu_src_ld = 0;
    // pinned staging + second stream for the host-buffer entry points
    void *pinned[2] = {nullptr, nullptr};
    size_t pinned_bytes = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bmb200_halo halo;
};

#define BMB_CUDA(h, call)                                                                   \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            snprintf((h)->err, sizeof((h)->err), "%s:%d %s: %s", __FILE__, __LINE__, #call, \
                     cudaGetErrorString(e_));                                               \
            return BMB200_ERR_CUDA - (int)e_;                                               \
        }                                                                                   \
    } while (0)

#define BMB_LAUNCH_CHECK(h)          \
    do {                             \
        (h)->launches++;             \
        BMB_CUDA(h, cudaGetLastError()); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int bmb_ensure_scratch(bmb200_ctx *h, size_t bytes);

static inline i64 imin64(i64 a, i64 b) { return a < b ? a : b; }
static inline i64 imax64(i64 a, i64 b) { return a > b ? a : b; }
static inline i64 cdiv64(i64 a, i64 b) { return (a + b - 1) / b; }
__device__ __forceinline__ i64 imin64_d(i64 a, i64 b) { return a < b ? a : b; }
__device__ __forceinline__ i64 imax64_d(i64 a, i64 b) { return a > b ? a : b; }

// ---- device helpers ---------------------------------------------------------------------------
__device__ __forceinline__ double ld_stream(const double *p) {  // read-once data: bypass L1 allocation
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void ld_stream_v2(const double *p, double &a, double &b) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "l"(p));
}
__device__ __forceinline__ void ld_stream_v4(const double *p, double &a, double &b, double &c, double &d) {
    // 256-bit global load (LDG.E.256, sm_100+)
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(a), "=d"(b), "=d"(c), "=d"(d)
                 : "l"(p));
}
__device__ __forceinline__ void st_stream(double *p, double v) {
    asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// ---- correctly rounded x / d from a precomputed correctly rounded reciprocal ---------------------------------------
// r = RN(1/d).  q0 = RN(x r) is within 2 ulp of x/d; one correction q1 = RN(q0 + (x - d q0) r) makes it faithful, and by
// Markstein's theorem a second one, q2 = RN(q1 + (x - d q1) r) with the remainder exact in an FMA, is RN(x/d) -- provided
// nothing over/underflows and d's significand is not all ones.  Those cases (and NaN/Inf/zero/subnormal operands) take
// the IEEE division instead, so the result is the true quotient bit for bit in every case.
static __device__ __noinline__ double gb_div_ieee(double x, double d) { return x / d; }
__device__ __forceinline__ bool gb_exp_mid(double v)  // 2^-500 <= |v| < 2^500 (excludes 0, subnormals, Inf, NaN)
{
    const unsigned e = ((unsigned)__double2hiint(v) >> 20) & 0x7ffu;
    return e - 523u <= 1000u;
}
__device__ __forceinline__ bool gb_div_safe_divisor(double d)
{
    const unsigned hi = (unsigned)__double2hiint(d) & 0xfffffu, lo = (unsigned)__double2loint(d);
    return gb_exp_mid(d) && !(hi == 0xfffffu && lo == 0xffffffffu);
}
__device__ __forceinline__ double gb_div(double x, double d, double r, bool dsafe)
{
    if (dsafe && gb_exp_mid(x)) {
        const double q0 = __dmul_rn(x, r);
        const double q1 = fma(fma(-q0, d, x), r, q0);
        return fma(fma(-q1, d, x), r, q1);
    }
    return gb_div_ieee(x, d);
}

