// Development-only: lets g++ compile the DEVICE math headers (prob3_device.cuh) so that numerics
// changes can be checked against the oracle in this GPU-less container before spending GPU time.
// Never part of the library (pisa_b200/build.py compiles with nvcc and without PISAB_HOST_EMU).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __constant__ static const
#define __restrict__
#define __launch_bounds__(...)
typedef int cudaError_t;
typedef void *cudaStream_t;
#define cudaSuccess 0
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }
template <typename T> static inline T __ldg(const T *p) { return *p; }
using std::fma; using std::fmax; using std::fmin; using std::rint; using std::sqrt;
// (sincos, exp, cbrt, hypot, atan2 of prob3_decay.cuh: glibc's, declared by <cmath> under _GNU_SOURCE)
static inline void pisab_emu_sincosf(float x, float *s, float *c) { *s = sinf(x); *c = cosf(x); }
#define __sincosf pisab_emu_sincosf
struct double2 { double x, y; };
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
#include <cstring>
static inline int __double2loint(double v) { uint64_t b; std::memcpy(&b, &v, 8); return (int)(uint32_t)(b & 0xffffffffull); }
static inline int __double2hiint(double v) { uint64_t b; std::memcpy(&b, &v, 8); return (int)(uint32_t)(b >> 32); }
static inline double __hiloint2double(int hi, int lo) {
    const uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double v; std::memcpy(&v, &b, 8); return v;
}
// the shared-memory state classes index by thread: never instantiated on the host, but they must parse
struct pisab_emu_dim3 { int x, y, z; };
static const pisab_emu_dim3 threadIdx = {0, 0, 0}, blockDim = {1, 1, 1}, blockIdx = {0, 0, 0}, gridDim = {1, 1, 1};
static inline double pisab_emu_hi32(double v) {
    uint64_t b; std::memcpy(&b, &v, 8); b &= 0xffffffff00000000ull; std::memcpy(&v, &b, 8); return v;
}
