// fp64_issue.cu -- how fast does ONE CTA issue FP64 work?  (the regime of the chain kernels: 1-8 warps on an SM)
//   per iteration and per thread: 16 independent DFMAs (a 4x4 register patch updated by a rank-1 product read from shared memory),
//   optionally a __syncthreads, with 32 / 256 / 1024 threads; also a dependent rsqrt chain and a 64-bit shuffle.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_issue tools/fp64_issue.cu
#include <cstdio>
#include <cuda_runtime.h>

template <bool BAR>
__global__ void patch_update(double *out, long long *cyc, int iters)
{
    __shared__ __align__(16) double xs[2][64];
    if (threadIdx.x < 64) { xs[0][threadIdx.x] = 1e-3 * threadIdx.x; xs[1][threadIdx.x] = 2e-3 * threadIdx.x; }
    __syncthreads();
    const int a = (threadIdx.x >> 4) & 15, b = threadIdx.x & 15;
    double v[4][4];
    for (int u = 0; u < 4; ++u) for (int w = 0; w < 4; ++w) v[u][w] = u + w;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const double *x = xs[it & 1];
        const double2 r01 = *reinterpret_cast<const double2 *>(x + 4 * a), r23 = *reinterpret_cast<const double2 *>(x + 4 * a + 2);
        const double2 c01 = *reinterpret_cast<const double2 *>(x + 4 * b), c23 = *reinterpret_cast<const double2 *>(x + 4 * b + 2);
        const double xr[4] = {r01.x, r01.y, r23.x, r23.y}, xc[4] = {c01.x, c01.y, c23.x, c23.y};
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int w = 0; w < 4; ++w) v[u][w] = fma(-xr[u], xc[w], v[u][w]);
        if (BAR) __syncthreads();
    }
    const long long t1 = clock64();
    double t = 0;
    for (int u = 0; u < 4; ++u) for (int w = 0; w < 4; ++w) t += v[u][w];
    out[threadIdx.x] = t;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

__global__ void rsqrt_chain(double *out, long long *cyc, int iters, double a0)
{
    double a = a0 + threadIdx.x;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
        const double h = 0.5 * a;
        y = y * fma(-(h * y), y, 1.5);
        y = y * fma(-(h * y), y, 1.5);
        a = a * y + 3.0;  // dependent
    }
    const long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

__global__ void sqrt_div_chain(double *out, long long *cyc, int iters, double a0)
{
    double a = a0 + threadIdx.x;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const double d = sqrt(a);
        a = 1.0 / d + 3.0;
    }
    const long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

__global__ void shfl_chain(double *out, long long *cyc, int iters)
{
    double a = threadIdx.x;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) a = __shfl_sync(0xffffffffu, a, (threadIdx.x + 1) & 31) + 1.0;
    const long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main()
{
    double *out;
    long long *cyc, h;
    cudaMalloc(&out, 1024 * sizeof(double));
    cudaMalloc(&cyc, sizeof(long long));
    const int iters = 2000;
    const int nts[3] = {32, 256, 1024};
    printf("{");
    for (int i = 0; i < 3; ++i) {
        patch_update<false><<<1, nts[i]>>>(out, cyc, iters);
        cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("\"patch16_nobar_%d\": %.1f, ", nts[i], (double)h / iters);
        patch_update<true><<<1, nts[i]>>>(out, cyc, iters);
        cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("\"patch16_bar_%d\": %.1f, ", nts[i], (double)h / iters);
    }
    rsqrt_chain<<<1, 32>>>(out, cyc, iters, 2.0);
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("\"rsqrt_newton2_plus_fma\": %.1f, ", (double)h / iters);
    sqrt_div_chain<<<1, 32>>>(out, cyc, iters, 2.0);
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("\"sqrt_div_dadd\": %.1f, ", (double)h / iters);
    shfl_chain<<<1, 32>>>(out, cyc, iters);
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("\"shfl64_dadd\": %.1f}\n", (double)h / iters);
    return 0;
}
