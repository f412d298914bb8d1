// dmma_lat.cu -- issue/latency behaviour of DMMA.8x8x4 (mma.sync.m8n8k4.f64): cycles per DMMA for one CTA of W warps (W/4 per
// SM sub-partition) with K independent accumulator chains per warp.  Decides how many chains gbmm needs in flight.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int K>
__global__ void dl(double *out, long long *cyc, int iters)
{
    double c[2 * K];
    for (int i = 0; i < 2 * K; ++i) c[i] = threadIdx.x * 1e-3 + i;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-3;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < K; ++i) dmma884(c[2 * i], c[2 * i + 1], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 2 * K; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main()
{
    double *o; long long *c, h;
    cudaMalloc(&o, 8 * 1024 * 148); cudaMalloc(&c, 8);
    const int iters = 4096;
    const int warps[] = {1, 4, 8, 16, 32};
    for (int wi = 0; wi < 5; ++wi) {
        const int W = warps[wi];
        printf("warps/SM %2d:", W);
#define RUN(K) { dl<K><<<148, 32 * W>>>(o, c, iters); dl<K><<<148, 32 * W>>>(o, c, iters); cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); \
        printf("  K=%d %.1f cyc/iter (%.1f per DMMA per SMSP-slot)", K, (double)h / iters, (double)h / iters / K / ((W + 3) / 4)); }
        RUN(1) RUN(2) RUN(3) RUN(4) RUN(8)
        printf("\n");
    }
    return 0;
}
