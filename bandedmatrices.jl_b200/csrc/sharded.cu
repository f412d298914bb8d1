// sharded.cu -- multi-GPU row-sharded gbmv with the x halo exchanged over NVLink peer memory.
//
// One process per GPU.  Rank r owns rows [c0,c1) of the square n x n band matrix.  Its band slab holds data
// columns [cs,ce) = [max(0,c0-kl), min(n,c1+ku)) -- the (kl+ku)-column data halo is static and replicated once at
// distribution time -- so the local product is the sub-matrix gbmv  y[c0:c1] = A[c0:c1, cs:ce] * x[cs:ce]  whose
// bandwidths in the same storage are (kl-(c0-cs), ku+(c0-cs)).  Only x[cs:c0) and x[c1:ce) live on the neighbours.
//
// The exchange is fused into the streaming kernel: warp 0 of block 0 first PUSHES this rank's boundary entries of x
// into the neighbours' mailboxes with peer stores (NVLink) and publishes an epoch flag (st.release.sys); every other
// warp starts streaming its interior columns immediately; only the two warps whose column sets touch a halo spin on
// their own mailbox flag (ld.acquire.sys) before reading it.  No NCCL call, no extra launch, no host sync: the
// ~2 us NVLink round trip hides behind ~200 us of interior work.  Mailboxes are double-buffered by epoch parity
// (a neighbour can be at most one call ahead, because its call e+1 needs our push e+1).
// Every row is accumulated in the same order as on one GPU: the sharded result is bit-identical.
#include "common.cuh"
#include "gbmv_systolic.cuh"

struct HaloBox {            // layout of one rank's mailbox (device memory, IPC-exported)
    // doubles: from_left[2][max_halo], from_right[2][max_halo]; then 4 x u64 flags {fl0, fl1, fr0, fr1}
    double *base;
    i64 max_halo;
    __host__ __device__ double *from_left(int par) const { return base + (i64)par * max_halo; }
    __host__ __device__ double *from_right(int par) const { return base + (2 + (i64)par) * max_halo; }
    __host__ __device__ unsigned long long *flags() const { return (unsigned long long *)(base + 4 * max_halo); }
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_volatile(const double *p)
{
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

struct XHalo {
    const double *x;    // local slice, S-column c lives at x[c - hl]
    const double *xl;   // own mailbox, entries pushed by the left neighbour  (S columns [0, hl))
    const double *xr;   // own mailbox, entries pushed by the right neighbour (S columns [hl+nl, hl+nl+hr))
    i64 hl, nl, hr;
    const unsigned long long *flag_l, *flag_r;
    unsigned long long epoch;
    int *err;
    __device__ __forceinline__ void wait(const unsigned long long *f) const
    {
        const long long t0 = clock64();
        while (ld_acquire_sys(f) < epoch) {
            if (clock64() - t0 > 6000000000LL) {  // ~3 s: a neighbour never pushed; fail loudly instead of hanging
                atomicExch(err, 1);
                break;
            }
        }
    }
    __device__ __forceinline__ void prepare(i64 set_base, int) const
    {
        if (hl > 0 && set_base < hl) wait(flag_l);
        if (hr > 0 && set_base + 32 > hl + nl) wait(flag_r);
    }
    __device__ __forceinline__ double stream(i64 c) const
    {
        if (c < hl) return ld_volatile(xl + c);
        if (c >= hl + nl) return ld_volatile(xr + (c - hl - nl));
        return ld_stream(x + (c - hl));
    }
    __device__ __forceinline__ double plain(i64 c) const
    {
        if (c < hl) { wait(flag_l); return ld_volatile(xl + c); }
        if (c >= hl + nl) { wait(flag_r); return ld_volatile(xr + (c - hl - nl)); }
        return x[c - hl];
    }
};

template <int W, int LDV>
__global__ void __launch_bounds__(256)
gbmv_n_systolic_sharded(i64 m, i64 n, int kl, int ku, double alpha, const double *__restrict__ a, i64 lda, XHalo xs,
                        double beta, double *__restrict__ y, i64 total_sets, i64 sets_per_run, i64 num_runs,
                        HaloBox left, HaloBox right, int push_l, int push_r, int par)
{
    // Programmatic dependent launch: the NEXT product in the stream may be launched now, so that its blocks take the SMs
    // this grid's blocks leave one by one (launch latency and block dispatch hide behind this grid's tail); its blocks
    // wait in griddepcontrol.wait -- before touching memory -- until this grid has completed and flushed.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (blockIdx.x == 0 && threadIdx.x < 32) {  // push first, then behave like any other warp
        const int lane = threadIdx.x;
        if (left.base && lane < push_l) left.from_right(par)[lane] = xs.x[lane];                     // my first ku entries
        if (right.base && lane < push_r) right.from_left(par)[lane] = xs.x[xs.nl - push_r + lane];   // my last kl entries
        __threadfence_system();
        __syncwarp();
        if (lane == 0) {
            if (left.base) st_release_sys(left.flags() + 2 + par, xs.epoch);   // the left rank's "from right" flag
            if (right.base) st_release_sys(right.flags() + par, xs.epoch);     // the right rank's "from left" flag
        }
    }
    gbmv_n_systolic_body<W, LDV, XHalo>(m, n, kl, ku, alpha, a, lda, xs, beta, y, total_sets, sets_per_run, num_runs);
}

extern "C" int bmb200_halo_destroy(bmb200_handle_t h)
{
    if (!h) return -1;
    DeviceGuard g(h->device);
    bmb200_halo &H = h->halo;
    if (H.left_box) cudaIpcCloseMemHandle(H.left_box);
    if (H.right_box) cudaIpcCloseMemHandle(H.right_box);
    if (H.box) cudaFree(H.box);
    H = bmb200_halo();
    return 0;
}

static size_t box_bytes(i64 max_halo) { return (size_t)(4 * max_halo) * sizeof(double) + 4 * sizeof(unsigned long long); }

extern "C" int bmb200_halo_create(bmb200_handle_t h, int64_t max_halo, void *ipc_handle_out)
{
    if (!h) return -1;
    if (max_halo < 1) return -2;
    if (!ipc_handle_out) return -3;
    DeviceGuard g(h->device);
    bmb200_halo_destroy(h);
    static_assert(sizeof(cudaIpcMemHandle_t) == BMB200_IPC_HANDLE_BYTES, "IPC handle size");
    BMB_CUDA(h, cudaMalloc(&h->halo.box, box_bytes(max_halo)));
    BMB_CUDA(h, cudaMemset(h->halo.box, 0, box_bytes(max_halo)));
    BMB_CUDA(h, cudaDeviceSynchronize());
    h->halo.max_halo = max_halo;
    cudaIpcMemHandle_t ih;
    BMB_CUDA(h, cudaIpcGetMemHandle(&ih, h->halo.box));
    memcpy(ipc_handle_out, &ih, sizeof(ih));
    return 0;
}

extern "C" int bmb200_halo_connect(bmb200_handle_t h, int rank, int nranks, const void *ipc_left, const void *ipc_right)
{
    if (!h) return -1;
    if (nranks < 1 || rank < 0 || rank >= nranks) return -2;
    if (!h->halo.box) return -1;
    DeviceGuard g(h->device);
    h->halo.rank = rank;
    h->halo.nranks = nranks;
    cudaIpcMemHandle_t ih;
    if (ipc_left) {
        memcpy(&ih, ipc_left, sizeof(ih));
        BMB_CUDA(h, cudaIpcOpenMemHandle((void **)&h->halo.left_box, ih, cudaIpcMemLazyEnablePeerAccess));
    }
    if (ipc_right) {
        memcpy(&ih, ipc_right, sizeof(ih));
        BMB_CUDA(h, cudaIpcOpenMemHandle((void **)&h->halo.right_box, ih, cudaIpcMemLazyEnablePeerAccess));
    }
    return 0;
}

template <int W, int LDV>
static int launch_sharded(bmb200_ctx *h, i64 ms, i64 ns, i64 kls, i64 kus, double alpha, const double *dA, i64 lda,
                          XHalo xs, double beta, double *dy, HaloBox L, HaloBox R, int push_l, int push_r, int par)
{
    const int threads = 256;
    static int per_sm_cached = 0;  // a property of the kernel image: queried once per instantiation, not per call
    if (per_sm_cached == 0)
        BMB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_cached, gbmv_n_systolic_sharded<W, LDV>, threads, 0));
    const int per_sm = per_sm_cached;
    const SystolicPlan p = systolic_plan(ms, kus, h->sm_count, per_sm, threads);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)p.blocks);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    BMB_CUDA(h, cudaLaunchKernelEx(&cfg, gbmv_n_systolic_sharded<W, LDV>, ms, ns, (int)kls, (int)kus, alpha, dA, lda, xs, beta, dy,
                                   p.total_sets, p.sets_per_run, p.num_runs, L, R, push_l, push_r, par));
    BMB_LAUNCH_CHECK(h);
    return 0;
}

#define SH_CASE(WW)                                                                                                    \
    case WW:                                                                                                           \
        return vec ? launch_sharded<WW, 8>(h, ms, ns, kls, kus, alpha, dA_local, lda, xs, beta, dy_local, L, R, push_l, \
                                           push_r, par)                                                                \
                   : launch_sharded<WW, 0>(h, ms, ns, kls, kus, alpha, dA_local, lda, xs, beta, dy_local, L, R, push_l, \
                                           push_r, par);

#define SH_CASE_S(WW)                                                                                                  \
    case WW:                                                                                                           \
        return launch_sharded<WW, 0>(h, ms, ns, kls, kus, alpha, dA_local, lda, xs, beta, dy_local, L, R, push_l, push_r, par);

extern "C" int bmb200_dgbmv_sharded(bmb200_handle_t h, int64_t n_global, int64_t c0, int64_t c1, int64_t kl,
                                    int64_t ku, double alpha, const double *dA_local, int64_t lda,
                                    const double *dx_local, double beta, double *dy_local)
{
    if (!h) return -1;
    if (n_global < 0) return -2;
    if (c0 < 0 || c0 > c1) return -3;
    if (c1 > n_global) return -4;
    if (kl < 0) return -5;
    if (ku < 0) return -6;
    if (lda < kl + ku + 1) return -9;
    bmb200_halo &H = h->halo;
    const i64 nl = c1 - c0;
    if (!H.box || H.max_halo < imax64(kl, ku)) {
        snprintf(h->err, sizeof(h->err), "dgbmv_sharded: call bmb200_halo_create/connect first (max_halo >= max(kl,ku))");
        return BMB200_ERR_CUDA;
    }
    if (nl < imax64(kl, ku) || kl + ku + 1 > 16) {
        snprintf(h->err, sizeof(h->err), "dgbmv_sharded: needs slab >= max(kl,ku) rows and kl+ku+1 <= 16 (got nl=%lld, band %lld)",
                 (long long)nl, (long long)(kl + ku + 1));
        return BMB200_ERR_CUDA;
    }
    DeviceGuard g(h->device);
    const i64 cs = imax64(0, c0 - kl), ce = imin64(n_global, c1 + ku);
    const i64 hl = c0 - cs, hr = ce - c1;
    const i64 ms = nl, ns = ce - cs, kus = ku + hl, kls = kl - hl;
    if ((hl > 0 && !H.left_box) || (hr > 0 && !H.right_box)) {
        snprintf(h->err, sizeof(h->err), "dgbmv_sharded: slab [%lld,%lld) needs a neighbour that is not connected",
                 (long long)c0, (long long)c1);
        return BMB200_ERR_CUDA;
    }
    H.epoch += 1;  // only once every validation has passed: a failed call must not desynchronise the ranks' epochs
    const int par = (int)(H.epoch & 1);
    HaloBox own{H.box, H.max_halo}, L{H.left_box, H.max_halo}, R{H.right_box, H.max_halo};
    XHalo xs;
    xs.x = dx_local;
    xs.xl = own.from_left(par);
    xs.xr = own.from_right(par);
    xs.hl = hl;
    xs.nl = nl;
    xs.hr = hr;
    xs.flag_l = own.flags() + par;
    xs.flag_r = own.flags() + 2 + par;
    xs.epoch = H.epoch;
    xs.err = h->d_info + 8;
    // what the neighbours need from me: the left rank reads my first ku entries, the right rank my last kl entries
    const int push_l = (H.left_box && c0 > 0) ? (int)ku : 0;
    const int push_r = (H.right_box && c1 < n_global) ? (int)kl : 0;
    // (alpha == 0 still runs the kernel: the epoch must be published or the neighbours would wait for this rank)
    const bool vec = (lda == 8) && (((uintptr_t)dA_local & 31) == 0);
    const int W = (int)(kl + ku + 1);
    switch (W) {
        SH_CASE(1) SH_CASE(2) SH_CASE(3) SH_CASE(4) SH_CASE(5) SH_CASE(6) SH_CASE(7) SH_CASE(8)
        SH_CASE_S(9) SH_CASE_S(10) SH_CASE_S(11) SH_CASE_S(12) SH_CASE_S(13) SH_CASE_S(14) SH_CASE_S(15) SH_CASE_S(16)
    default: break;
    }
    return -5;
}
