// Measures the FP64 roofline denominators on this B200 (MEASURED_PEAKS.json has no FP64 entry):
//   (1) register-resident DFMA loop, (2) register-resident DMMA.8x8x4 loop,
// and checks whether DMMA.8x8x4 equals the sequential FMA chain k = 0..3 bit for bit.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void dfma_loop(double *out, int iters, double s)
{
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], s, 1e-9);
    }
    double t = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

__global__ void dmma_loop(double *out, int iters, double s)
{
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = 0.0;
    double a = s * (threadIdx.x & 3), b = s * (threadIdx.x >> 2);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(c[2 * i], c[2 * i + 1], a, b);
    }
    double t = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

// one warp: D = A(8x4) * B(4x8) + C with DMMA, and the same with a sequential FMA chain
__global__ void dmma_exact(const double *A, const double *B, const double *C, double *Dm, double *Df)
{
    const int l = threadIdx.x;
    double a = A[(l >> 2) * 4 + (l & 3)];       // row-major A[row=l/4][k=l%4]
    double b = B[(l & 3) * 8 + (l >> 2)];       // B[k=l%4][n=l/4]
    const int row = l >> 2, col = 2 * (l & 3);
    double c0 = C[row * 8 + col], c1 = C[row * 8 + col + 1];
    double d0 = c0, d1 = c1;
    dmma884(d0, d1, a, b);
    Dm[row * 8 + col] = d0;
    Dm[row * 8 + col + 1] = d1;
    double f0 = c0, f1 = c1;
    for (int k = 0; k < 4; ++k) {
        f0 = fma(A[row * 4 + k], B[k * 8 + col], f0);
        f1 = fma(A[row * 4 + k], B[k * 8 + col + 1], f1);
    }
    Df[row * 8 + col] = f0;
    Df[row * 8 + col + 1] = f1;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 4, threads = 256, iters = 20000;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms;
    double best_fma = 0, best_mma = 0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        dfma_loop<<<blocks, threads>>>(out, iters, 0.999999);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double tf = 2.0 * 16 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
        if (tf > best_fma) best_fma = tf;
        cudaEventRecord(e0);
        dmma_loop<<<blocks, threads>>>(out, iters, 1e-3);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        tf = 2.0 * 256 * 8 * iters * (double)blocks * (threads / 32) / (ms * 1e-3) / 1e12;
        if (tf > best_mma) best_mma = tf;
    }
    // exactness
    double hA[32], hB[32], hC[64], hDm[64], hDf[64];
    srand(7);
    int mism = 0, trials = 2000;
    double *dA, *dB, *dC, *dDm, *dDf;
    cudaMalloc(&dA, 256); cudaMalloc(&dB, 256); cudaMalloc(&dC, 512); cudaMalloc(&dDm, 512); cudaMalloc(&dDf, 512);
    for (int t = 0; t < trials; ++t) {
        for (int i = 0; i < 32; ++i) { hA[i] = rand() / (double)RAND_MAX - 0.5; hB[i] = rand() / (double)RAND_MAX - 0.5; }
        for (int i = 0; i < 64; ++i) hC[i] = rand() / (double)RAND_MAX - 0.5;
        cudaMemcpy(dA, hA, 256, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, 256, cudaMemcpyHostToDevice);
        cudaMemcpy(dC, hC, 512, cudaMemcpyHostToDevice);
        dmma_exact<<<1, 32>>>(dA, dB, dC, dDm, dDf);
        cudaMemcpy(hDm, dDm, 512, cudaMemcpyDeviceToHost); cudaMemcpy(hDf, dDf, 512, cudaMemcpyDeviceToHost);
        for (int i = 0; i < 64; ++i) if (hDm[i] != hDf[i]) ++mism;
    }
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.2f, \"dmma_8x8x4_tflops\": %.2f, "
           "\"dmma_equals_sequential_fma_chain\": %s, \"dmma_mismatching_elements\": %d, \"dmma_elements_checked\": %d}\n",
           p.name, p.multiProcessorCount, best_fma, best_mma, mism == 0 ? "true" : "false", mism, trials * 64);
    return 0;
}
