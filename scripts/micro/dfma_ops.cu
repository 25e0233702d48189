// Microbenchmark: FP64 pipe throughput by operand pattern (all 16 warps/SM... 4 CTAs x 512 threads per SM, ILP 8).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(int iters, double a, double b, double *out)
{
    double x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x + i; y[i] = 1.0 + 1e-9 * (threadIdx.x + i); }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) x[i] = fma(x[i], a, b);                       // one register source + two uniform constants
            if (MODE == 1) x[i] = fma(x[i], y[i], b);                    // two register sources
            if (MODE == 2) x[i] = fma(x[i], y[i], y[(i + 3) & 7]);       // three distinct register sources
            if (MODE == 3) x[i] = x[i] * y[i];                           // DMUL
            if (MODE == 4) x[i] = x[i] + y[i];                           // DADD
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i] + y[i];
    if (s == 12345.678) out[0] = s;
}
template <int MODE> void run(const char *name)
{
    double *out; cudaMalloc(&out, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000, grid = 148 * 4, block = 512;
    k<MODE><<<grid, block>>>(iters, 0.999999, 1e-9, out);
    cudaEventRecord(e0);
    k<MODE><<<grid, block>>>(iters, 0.999999, 1e-9, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-40s %.2f T thread-instr/s\n", name, (double)grid * block * 8.0 * iters / (ms * 1e-3) / 1e12);
    cudaFree(out);
}
int main()
{
    run<0>("DFMA reg*const+const"); run<1>("DFMA reg*reg+const"); run<2>("DFMA reg*reg+reg (3 distinct)"); run<3>("DMUL reg*reg"); run<4>("DADD reg+reg");
    return 0;
}
