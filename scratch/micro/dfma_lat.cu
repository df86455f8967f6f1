// DFMA latency / throughput vs parallelism on B200 (development micro-benchmark).
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(int iters, double seed, double *sink, long long *cycles) {
    double a[CH];
    for (int c = 0; c < CH; ++c) a[c] = seed + threadIdx.x + c;
    const double m = 0.999999, b = 1e-7;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CH; ++c) a[c] = fma(a[c], m, b);
    }
    long long t1 = clock64();
    double s = 0; for (int c = 0; c < CH; ++c) s += a[c];
    if (s == 12345.678) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}
template <int CH> void run(int warps_per_sm) {
    double *sink; long long *cyc; cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
    const int iters = 20000;
    k<CH><<<148, warps_per_sm * 32>>>(iters, 1.0, sink, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    // per SMSP: warps_per_sm/4 warps (if >= 4), each issuing CH*iters DFMA
    double per_smsp_warps = warps_per_sm / 4.0;
    printf("chains %d warps/SM %2d: %.2f cycles per DFMA per warp-chain step; SMSP DFMA/cycle %.3f (peak 0.5)\n", CH, warps_per_sm,
           (double)c / iters / 1.0, per_smsp_warps * CH * iters / (double)c);
    cudaFree(sink); cudaFree(cyc);
}
int main() {
    for (int w : {4, 8, 16, 32}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
    return 0;
}
