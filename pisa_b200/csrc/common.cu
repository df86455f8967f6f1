// common.cu -- library plumbing: error text, launch accounting, device info, FP64 peak probe.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace pisab {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_profiling{0};
static thread_local double g_last_ms = -1.0;

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return PISAB_ERR_CUDA;
}

void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// arithmetic of the *_f32 propagation entry points (pisab_set_f32_math)
static std::atomic<int> g_f32_math{PISAB_F32_MATH_MIXED};
bool f32_math_mixed() { return g_f32_math.load(std::memory_order_relaxed) == PISAB_F32_MATH_MIXED; }

// CUDA-event timing of one launch on the stream it is issued to (bench.py roofline numbers).
static thread_local cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
LaunchTimer::LaunchTimer(cudaStream_t s) : stream(s), active(g_profiling.load() != 0) {
    if (!active) return;
    if (!g_ev0) {
        cudaEventCreate(&g_ev0);
        cudaEventCreate(&g_ev1);
    }
    cudaEventRecord(g_ev0, stream);
}
LaunchTimer::~LaunchTimer() {
    if (!active) return;
    cudaEventRecord(g_ev1, stream);
    cudaEventSynchronize(g_ev1);
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, g_ev0, g_ev1) == cudaSuccess) g_last_ms = ms;
}

int sm_count() {
    static int cached = -1;
    if (cached >= 0) return cached;
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    cached = n;
    return n;
}

// ---- FP64 roofline denominator: 8 independent DFMA chains per thread, full grid ------------
__global__ void __launch_bounds__(256) dfma_probe_kernel(int iters, double seed, double *sink) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-7;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 123.456) sink[0] = r; // never true; keeps the chains alive
}

} // namespace pisab

using namespace pisab;

extern "C" {

const char *pisab_last_error(void) { return g_err; }
const char *pisab_version(void) { return "pisa_b200 0.1 (sm_100a)"; }

int pisab_device_info(int32_t *sms, int32_t *cc_major, int32_t *cc_minor) {
    int dev = 0;
    PISAB_CUDA_CHECK(cudaGetDevice(&dev));
    cudaDeviceProp p;
    PISAB_CUDA_CHECK(cudaGetDeviceProperties(&p, dev));
    if (sms) *sms = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return PISAB_OK;
}

int pisab_set_f32_math(int32_t mode) {
    if (mode != PISAB_F32_MATH_FP64 && mode != PISAB_F32_MATH_MIXED) {
        set_error("f32 math mode must be PISAB_F32_MATH_FP64 (0) or PISAB_F32_MATH_MIXED (1)");
        return PISAB_ERR_ARG;
    }
    g_f32_math.store(mode);
    return PISAB_OK;
}
int pisab_get_f32_math(void) { return g_f32_math.load(); }

int64_t pisab_launch_count(int32_t reset) {
    const long long v = g_launches.load();
    if (reset) g_launches.store(0);
    return v;
}

int pisab_set_profiling(int32_t on) {
    g_profiling.store(on ? 1 : 0);
    return PISAB_OK;
}

double pisab_last_kernel_ms(void) { return g_last_ms; }

int pisab_fp64_peak_probe(int32_t iters, double *flops_per_s, double *elapsed_ms) {
    if (iters < 1) { set_error("iters < 1"); return PISAB_ERR_ARG; }
    const int sms = sm_count();
    if (sms <= 0) { set_error("no CUDA device"); return PISAB_ERR_CUDA; }
    double *sink = nullptr;
    PISAB_CUDA_CHECK(cudaMalloc(&sink, 8));
    const int blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    PISAB_CUDA_CHECK(cudaEventCreate(&e0));
    PISAB_CUDA_CHECK(cudaEventCreate(&e1));
    dfma_probe_kernel<<<blocks, threads>>>(iters / 8 + 1, 1.0, sink); // warm-up
    note_launch();
    PISAB_CUDA_CHECK(cudaEventRecord(e0, 0));
    dfma_probe_kernel<<<blocks, threads>>>(iters, 1.0, sink);
    note_launch();
    PISAB_CUDA_CHECK(cudaEventRecord(e1, 0));
    PISAB_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    PISAB_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    const double flop = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
    if (flops_per_s) *flops_per_s = flop / (ms * 1e-3);
    if (elapsed_ms) *elapsed_ms = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return PISAB_OK;
}

} // extern "C"
