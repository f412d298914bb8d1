// pingpong.cu -- cost of handing a 16-double panel from one warp to the next through shared memory (same SM), the
// hand-over that sits on the dependency chain of gbtrs_cluster.cu / gbtrf_pipe.cu.  NW warps pass a token round-robin.
//   mode 0: data STS, fence.acq_rel.cta, tag STS; consumer polls the tag (ld.volatile.shared), fence, 8 x LDS.128
//   mode 1: same without fences
//   mode 2: mbarrier: producer STS + mbarrier.arrive (release.cta), consumer mbarrier.try_wait.parity
#include <cstdio>
#include <cuda_runtime.h>
#define HOPS 8192
template <int MODE>
__global__ void pp(long long *cyc, double *out, int nw)
{
    __shared__ __align__(16) double data[32][16];
    __shared__ unsigned tag[32];
    __shared__ unsigned long long bar[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x < 32) { tag[threadIdx.x] = 0; if (MODE == 2) { unsigned a = (unsigned)__cvta_generic_to_shared(&bar[threadIdx.x]); asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"(a)); } }
    __syncthreads();
    double x = 1.0 + lane * 1e-9;
    long long t0 = clock64();
    for (int h = w; h < HOPS; h += nw) {
        // wait for hop h-1 (produced by warp w-1) unless first
        if (h > 0) {
            const int src = (h - 1) % nw;
            if (MODE == 2) {
                const unsigned a = (unsigned)__cvta_generic_to_shared(&bar[src]);
                const unsigned par = ((h - 1) / nw) & 1;
                unsigned done = 0;
                while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(a), "r"(par) : "memory");
            } else {
                const unsigned ta = (unsigned)__cvta_generic_to_shared(&tag[src]);
                unsigned tg;
                do { asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(tg) : "r"(ta) : "memory"); } while (tg != (unsigned)h);
                if (MODE == 0) asm volatile("fence.acq_rel.cta;" ::: "memory");
            }
            const unsigned ra = (unsigned)__cvta_generic_to_shared(&data[src][0]);
            double v[16];
#pragma unroll
            for (int c = 0; c < 16; c += 2) asm volatile("ld.volatile.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[c]), "=d"(v[c + 1]) : "r"(ra + 8u * c) : "memory");
#pragma unroll
            for (int c = 0; c < 16; ++c) x += v[c];  // 16 dependent DADDs stand in for the 16 DFMAs of a near step
        }
        // publish hop h
        if (lane < 16) data[w][lane] = x;
        if (MODE == 2) {
            __syncwarp();
            if (lane == 0) { const unsigned a = (unsigned)__cvta_generic_to_shared(&bar[w]); unsigned long long st; asm volatile("mbarrier.arrive.release.cta.shared.b64 %0, [%1];" : "=l"(st) : "r"(a) : "memory"); }
        } else {
            if (MODE == 0) asm volatile("fence.acq_rel.cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) *(volatile unsigned *)&tag[w] = (unsigned)(h + 1);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = x;
}
int main()
{
    long long *c, h; double *o;
    cudaMalloc(&c, 8); cudaMalloc(&o, 1024 * 8);
    const int nws[] = {2, 4, 7, 8};
    for (int mode = 0; mode < 3; ++mode)
        for (int k = 0; k < 4; ++k) {
            const int nw = nws[k];
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) pp<0><<<1, 32 * nw>>>(c, o, nw);
                if (mode == 1) pp<1><<<1, 32 * nw>>>(c, o, nw);
                if (mode == 2) pp<2><<<1, 32 * nw>>>(c, o, nw);
            }
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
            printf("mode %d warps %d: %.1f cycles per hop (incl. 16 dependent DADD ~ 130)  %s\n", mode, nw, (double)h / HOPS, cudaGetErrorString(e));
        }
    return 0;
}
