// Microbenchmark: the plain-data round of mh_likelihood<4> (4 proposals x 2 logarithms per thread, operands from shared
// memory) in isolation, for 1..6 warps per SM sub-partition, to see what limits a warp: its own dependency latencies (time per
// round does not change with co-resident warps), the FP64 pipe (22 FP64 warp-instructions x 2 cycles per term) or something else.
// With the fdlibm-style logarithm of round 1 (16 FP64 + MUFU.RCP64H each) this loop saturated at 72 % of the FP64 pipe however
// many warps ran: the reciprocal seeds and conversions on the XU pipe were a second limiter.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I phylowgs_b200/csrc -o /tmp/plain_round scripts/micro/plain_round.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "pwgs_device.cuh"

#ifndef BU
#define BU 4
#endif
template <int REGS>
__global__ void __maxnreg__(REGS) k(int rounds, int K, long long *cyc, double *out)
{
    extern __shared__ __align__(16) double sm_all[];
    double *sm = sm_all + PWGS_LOG_TAB_DOUBLES;
    pwgs_log_tab_fill(sm_all, threadIdx.x, blockDim.x);
    const double2 *ltab = (const double2 *)sm_all + (threadIdx.x & (PWGS_LOG_TAB_COPIES - 1));
    double *phi = sm;                               // [BU][K]
    double *mur = sm + BU * K, *muv = mur + 1024;   // [1024] each
    int *node = (int *)(muv + 1024), *pa = node + 1024, *pd = pa + 1024;
    for (int i = threadIdx.x; i < BU * K; i += blockDim.x) phi[i] = 0.05 + 0.9 * ((i * 37) % 101) / 101.0;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) { mur[i] = 0.999; muv[i] = 0.499; node[i] = i % K; pa[i] = 30 + i % 17; pd[i] = 60 + i % 29; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc[BU];
#pragma unroll
    for (int b = 0; b < BU; b++) acc[b] = 0.0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < rounds; r++) {
        const int jj = ((r + warp * 7) * 32 + lane) & 1023;
        const double mu_r = mur[jj], dmu = muv[jj] - mu_r;
        const int nd = node[jj];
        const int a = pa[jj], d = pd[jj];
        const double fa = (double)a, fb = (double)(d - a);
        double xx[2 * BU], ll[2 * BU];
#pragma unroll
        for (int b = 0; b < BU; b++) {
            const double ph = phi[b * K + nd];
            xx[b] = fma(ph, dmu, mu_r);
            xx[BU + b] = 1.0 - xx[b];
        }
        pwgs_log_pos_n<2 * BU>(xx, ll, ltab);
#pragma unroll
        for (int b = 0; b < BU; b++) acc[b] = fma(fa, ll[b], fma(fb, ll[BU + b], acc[b]));
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int b = 0; b < BU; b++) s += acc[b];
    if (s == 12345.678) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int REGS>
void run(int warps)
{
    long long *cyc, h;
    double *out;
    cudaMalloc(&cyc, 8); cudaMalloc(&out, 8);
    const int K = 11, rounds = 4000;
    const size_t smem = (PWGS_LOG_TAB_DOUBLES + BU * K + 2048) * 8 + 3 * 1024 * 4;
    for (int rep = 0; rep < 2; rep++) k<REGS><<<148, warps * 32, smem>>>(rounds, K, cyc, out);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_round = (double)h / rounds;
    const double fp64_per_round = 22.0 * BU;
    printf("BU=%d regs<=%3d warps/SM=%2d (%.2f per sub-partition): %7.1f cycles per round per warp; FP64 pipe %.1f %% (%.0f warp-instr x 2 cycles x %.2f warps / round time)\n",
           BU, REGS, warps, warps / 4.0, per_round, 100.0 * fp64_per_round * 2.0 * (warps / 4.0) / per_round, fp64_per_round, warps / 4.0);
    cudaFree(cyc); cudaFree(out);
}

int main(int argc, char **argv)
{
    if (argc > 1) { run<168>(atoi(argv[1])); return 0; }      // one configuration (for ncu)
    for (int w : {1, 4, 8, 12}) run<168>(w);
    run<152>(13);
    run<128>(16);
    run<96>(20);
    run<80>(24);
    return 0;
}
