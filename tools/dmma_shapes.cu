// FP64 mma.sync shapes beyond m8n8k4 on this GPU (sm_90+ PTX: m16n8k4 / m16n8k8 / m16n8k16):
//   (1) does the instruction compile for sm_100a and produce the documented fragment layout,
//   (2) is its k-accumulation the SEQUENTIAL ascending FMA chain (bit for bit) -- the property the banded product relies on,
//   (3) what does a register-resident loop of it sustain (TFLOP/s), compared with DMMA.8x8x4.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_shapes tools/dmma_shapes.cu ; prints one JSON line
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>

template <int K>
struct Frag;
template <>
struct Frag<4> { static constexpr int NA = 2, NB = 1; };
template <>
struct Frag<8> { static constexpr int NA = 4, NB = 2; };
template <>
struct Frag<16> { static constexpr int NA = 8, NB = 4; };

__device__ __forceinline__ void mma16(double (&d)[4], const double (&a)[2], const double (&b)[1])
{
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
}
__device__ __forceinline__ void mma16(double (&d)[4], const double (&a)[4], const double (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16(double (&d)[4], const double (&a)[8], const double (&b)[4])
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// one warp: D = A(16xK) * B(Kx8) + C through the instruction, and through a sequential FMA chain k = 0..K-1
template <int K>
__global__ void exact(const double *A, const double *B, const double *C, double *Dm, double *Df)
{
    const int l = threadIdx.x, g = l >> 2, t = l & 3;
    double a[Frag<K>::NA], b[Frag<K>::NB], d[4];
#pragma unroll
    for (int i = 0; i < Frag<K>::NA; ++i) a[i] = A[(g + 8 * (i & 1)) * K + (t + 4 * (i >> 1))];  // row-major A[row][k]
#pragma unroll
    for (int i = 0; i < Frag<K>::NB; ++i) b[i] = B[(t + 4 * i) * 8 + g];                          // B[k][n]
    const int rows[4] = {g, g, g + 8, g + 8}, cols[4] = {2 * t, 2 * t + 1, 2 * t, 2 * t + 1};
#pragma unroll
    for (int i = 0; i < 4; ++i) d[i] = C[rows[i] * 8 + cols[i]];
    mma16(d, a, b);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        Dm[rows[i] * 8 + cols[i]] = d[i];
        double f = C[rows[i] * 8 + cols[i]];
        for (int k = 0; k < K; ++k) f = fma(A[rows[i] * K + k], B[k * 8 + cols[i]], f);
        Df[rows[i] * 8 + cols[i]] = f;
    }
}

template <int K>
__global__ void loop(double *out, int iters, double s)
{
    double c[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) c[j][i] = 0.0;
    double a[Frag<K>::NA], b[Frag<K>::NB];
#pragma unroll
    for (int i = 0; i < Frag<K>::NA; ++i) a[i] = s * ((threadIdx.x & 3) + i);
#pragma unroll
    for (int i = 0; i < Frag<K>::NB; ++i) b[i] = s * ((threadIdx.x >> 2) + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j) mma16(c[j], a, b);
    }
    double t = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) t += c[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int K>
static void run(const char *name, int sms)
{
    const int blocks = sms * 4, threads = 256, iters = 4000;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms;
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        loop<K><<<blocks, threads>>>(out, iters, 1e-3);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * 16 * 8 * K * 4 * iters * (double)blocks * (threads / 32) / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    double hA[16 * 16], hB[16 * 8], hC[128], hDm[128], hDf[128];
    double *dA, *dB, *dC, *dDm, *dDf;
    cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dC, sizeof(hC)); cudaMalloc(&dDm, sizeof(hDm)); cudaMalloc(&dDf, sizeof(hDf));
    int mism = 0, wrong = 0, trials = 1000;
    srand(11);
    for (int tr = 0; tr < trials; ++tr) {
        for (int i = 0; i < 16 * K; ++i) hA[i] = rand() / (double)RAND_MAX - 0.5;
        for (int i = 0; i < K * 8; ++i) hB[i] = rand() / (double)RAND_MAX - 0.5;
        for (int i = 0; i < 128; ++i) hC[i] = rand() / (double)RAND_MAX - 0.5;
        cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
        cudaMemcpy(dC, hC, sizeof(hC), cudaMemcpyHostToDevice);
        exact<K><<<1, 32>>>(dA, dB, dC, dDm, dDf);
        cudaMemcpy(hDm, dDm, sizeof(hDm), cudaMemcpyDeviceToHost); cudaMemcpy(hDf, dDf, sizeof(hDf), cudaMemcpyDeviceToHost);
        for (int i = 0; i < 128; ++i) {
            if (hDm[i] != hDf[i]) ++mism;
            if (fabs(hDm[i] - hDf[i]) > 1e-12) ++wrong;  // a layout error shows up here, a different summation order only above
        }
    }
    cudaError_t e = cudaGetLastError();
    printf("\"%s\": {\"tflops\": %.2f, \"layout_errors\": %d, \"elements_differing_from_sequential_fma_chain\": %d, \"elements_checked\": %d, \"cuda\": \"%s\"}",
           name, best, wrong, mism, trials * 128, cudaGetErrorString(e));
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("{\"gpu\": \"%s\", ", p.name);
    run<4>("m16n8k4", p.multiProcessorCount);
    printf(", ");
    run<8>("m16n8k8", p.multiProcessorCount);
    printf(", ");
    run<16>("m16n8k16", p.multiProcessorCount);
    printf("}\n");
    return 0;
}
