// Issue-rate micro-benchmark for the FP32-mode design (development tool, B200):
// warp-instructions per cycle per SM sub-partition for DFMA, scalar FFMA (3-register form), packed FFMA2
// (fma.rn.f32x2) and 1:1 / 1:2 mixes of them, with enough independent chains and warps to be throughput bound.
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float ffma(float a, float b, float c) {
    float r;
    asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ double dfma(double a, double b, double c) {
    double r;
    asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(r) : "d"(a), "d"(b), "d"(c));
    return r;
}
__device__ __forceinline__ int iadd3(int a, int b, int c) {
    int r;
    asm volatile("{ .reg .s32 t; add.s32 t, %1, %2; add.s32 %0, t, %3; }" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// MODE: 0 DFMA, 1 FFMA, 2 FFMA2, 3 DFMA+FFMA2 (1:1), 4 DFMA+FFMA (1:1), 5 DFMA + 2 FFMA2, 6 FFMA+FFMA2, 7 DFMA+IADD,
//       8 FFMA2 + IADD
template <int MODE, int CH>
__global__ void k(int iters, double seed, double *sink, long long *cycles) {
    double d[CH];
    float f[CH];
    u64 p[CH], p2[CH];
    int n[CH];
    for (int c = 0; c < CH; ++c) {
        d[c] = seed + threadIdx.x + c;
        f[c] = (float)d[c];
        p[c] = (u64)__float_as_uint(f[c]) | ((u64)__float_as_uint(f[c] + 1.f) << 32);
        p2[c] = p[c] + 1;
        n[c] = threadIdx.x + c;
    }
    const double dm = 0.999999, db = 1e-7;
    const float fm = 0.999999f, fb = 1e-7f;
    const u64 pm = (u64)__float_as_uint(fm) | ((u64)__float_as_uint(fm) << 32);
    const u64 pb = (u64)__float_as_uint(fb) | ((u64)__float_as_uint(fb) << 32);
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            if (MODE == 0 || MODE == 3 || MODE == 4 || MODE == 5 || MODE == 7) d[c] = dfma(d[c], dm, db);
            if (MODE == 1 || MODE == 4 || MODE == 6) f[c] = ffma(f[c], fm, fb);
            if (MODE == 2 || MODE == 3 || MODE == 5 || MODE == 6 || MODE == 8) p[c] = ffma2(p[c], pm, pb);
            if (MODE == 5) p2[c] = ffma2(p2[c], pm, pb);
            if (MODE == 7 || MODE == 8) n[c] = iadd3(n[c], n[c], i);
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int c = 0; c < CH; ++c) s += d[c] + f[c] + (double)(p[c] & 0xffff) + (double)(p2[c] & 0xff) + n[c];
    if (s == 12345.678) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int MODE, int CH>
void run(const char *name, int instr_per_step, int warps_per_sm) {
    double *sink;
    long long *cyc;
    cudaMalloc(&sink, 8);
    cudaMalloc(&cyc, 8);
    const int iters = 4000;
    k<MODE, CH><<<148, warps_per_sm * 32>>>(iters, 1.0, sink, cyc);
    cudaDeviceSynchronize();
    k<MODE, CH><<<148, warps_per_sm * 32>>>(iters, 1.0, sink, cyc);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_smsp = warps_per_sm / 4.0 * CH * iters * instr_per_step / (double)c;
    printf("%-22s chains %d warps/SM %2d: %.3f warp-instr/cycle/SMSP  (%.1f cycles per chain step)\n", name, CH,
           warps_per_sm, per_smsp, (double)c / iters);
    cudaFree(sink);
    cudaFree(cyc);
}

int main() {
    for (int w : {8, 16}) {
        run<0, 8>("DFMA", 1, w);
        run<1, 8>("FFMA", 1, w);
        run<2, 8>("FFMA2", 1, w);
        run<3, 8>("DFMA+FFMA2", 2, w);
        run<4, 8>("DFMA+FFMA", 2, w);
        run<5, 8>("DFMA+2xFFMA2", 3, w);
        run<6, 8>("FFMA+FFMA2", 2, w);
        run<7, 8>("DFMA+2xIADD", 3, w);
        run<8, 8>("FFMA2+2xIADD", 3, w);
    }
    return 0;
}
