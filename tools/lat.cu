// lat.cu -- dependent-chain latencies (cycles) of the instructions that sit on the LU / solve dependency chains.
// Single warp, one CTA; each test is a chain of N dependent operations timed with clock64().
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
__global__ void lat(double *out, long long *cyc, double seed, int src)
{
    __shared__ double sm[64];
    const int lane = threadIdx.x;
    sm[lane] = seed + lane; sm[lane + 32] = seed;
    __syncwarp();
    double x = seed + lane * 1e-9, y = 1.0 + 1e-12, z = 1e-13;
    long long t0, t1;
    int k = 0;
#define TIME(name, body)                                   \
    t0 = clock64();                                        \
    _Pragma("unroll 16") for (int i = 0; i < N; ++i) { body; } \
    t1 = clock64();                                        \
    if (lane == 0) cyc[k] = t1 - t0;                       \
    ++k;
    TIME(dfma, x = fma(x, y, z))
    TIME(dmul, x = x * y)
    TIME(dadd, x = x + z)
    TIME(shfl64, x = __shfl_sync(0xffffffffu, x, (lane + src) & 31))
    TIME(shfl_dfma, x = fma(__shfl_sync(0xffffffffu, x, (lane + src) & 31), y, z))
    TIME(fsel64, x = (x > 0.5) ? x : y)
    TIME(ddiv, x = y / x + 1.0)
    TIME(lds_chain, x = sm[((int)x) & 31] + 0.0 * x)
    TIME(shfl_fsel_dfma_fsel, { double t = __shfl_sync(0xffffffffu, x, (lane + src) & 31); double u = (lane == src) ? t : x; u = fma(u, y, z); x = (lane == 3) ? z : u; })
    TIME(redux, { unsigned m = __reduce_max_sync(0xffffffffu, (unsigned)__double2hiint(x)); x = x + (double)(m & 1); })
    TIME(ballot, { unsigned m = __ballot_sync(0xffffffffu, x > 0.3); x = x + (double)(m & 1); })
    TIME(sts_lds, { sm[lane] = x; __syncwarp(); x = sm[(lane + src) & 31]; __syncwarp(); })
    TIME(rcp64h, { asm volatile("{.reg .b32 lo, hi, r; mov.b64 {lo,hi}, %0; rcp.approx.ftz.f64 %0, %0; }" : "+d"(x)); x = x + 1.0; })
    out[lane] = x;
}
int main()
{
    double *o; long long *c, h[32];
    cudaMalloc(&o, 32 * 8); cudaMalloc(&c, 32 * 8);
    cudaMemset(c, 0, 32 * 8);
    for (int rep = 0; rep < 2; ++rep) lat<<<1, 32>>>(o, c, 0.75, 1);
    cudaMemcpy(h, c, 32 * 8, cudaMemcpyDeviceToHost);
    const char *names[] = {"dfma", "dmul", "dadd", "shfl64", "shfl64+dfma", "fsel64(+cmp)", "ddiv+dadd", "lds+dfma", "shfl+fsel+dfma+fsel", "redux+cvt+dadd", "ballot+cvt+dadd", "sts+sync+lds+sync", "rcp64h+dadd"};
    printf("{");
    for (int i = 0; i < 13; ++i) printf("\"%s\": %.1f%s", names[i], (double)h[i] / N, i < 12 ? ", " : "}\n");
    return 0;
}
