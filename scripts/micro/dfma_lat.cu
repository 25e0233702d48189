// Microbenchmark: FP64 FMA dependent-issue latency and per-SM throughput vs (warps per SM, independent chains per thread).
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(int iters, double a, double b, double *out, long long *cyc)
{
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    if (s == 12345.678) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP>
void run(int warps)
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    int iters = 20000;
    k<ILP><<<1, warps * 32>>>(iters, 0.999999, 1e-9, out, cyc);
    k<ILP><<<1, warps * 32>>>(iters, 0.999999, 1e-9, out, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per_instr = (double)h / ((double)iters * ILP);
    printf("warps/SM=%2d ILP=%d : %.2f cycles per DFMA per warp -> %.2f warp-DFMA/cycle/SM (peak 2.0)\n", warps, ILP, per_instr, warps / per_instr);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    for (int w : {1, 4, 8, 12, 16, 32}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
    return 0;
}
