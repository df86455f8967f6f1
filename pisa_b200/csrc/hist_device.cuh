// hist_device.cuh -- deterministic, privatised weighted histogram building blocks.
//
// Reference semantics (hist.py:198-209 -> translation.histogram -> fast_histogram.histogramdd):
//     hist[idx] += w ; sumw2[idx] += w*w        for every event whose index is in range.
// B200 design: every WARP owns a private copy of the bins (w and w^2) in shared memory.  One
// warp-collective step bins 32 events: lanes that hit the same bin are grouped with
// __match_any_sync, the lowest lane of each group sums its group's weights in ascending lane
// order and adds the result to the warp's private bin -- no atomics, so the accumulation order
// (and therefore every rounding) is fixed by (grid, block) alone.  Blocks write their partial
// histograms to a workspace and a second kernel reduces them in block order.
#pragma once
#include "common.cuh"

namespace pisab {

#ifndef PISAB_HIST_BLOCK
#define PISAB_HIST_BLOCK 128
#endif
constexpr int kHistBlock = PISAB_HIST_BLOCK;

struct WarpHist {
    double *bins;  // this warp's [2][n_bins]: w then w^2
    double *stage; // this warp's [32] staging slots
    double *all;   // block base
    int n_bins;

    __host__ __device__ static size_t smem_bytes(int block_threads, int n_bins) {
        return (size_t)(block_threads / 32) * (2 * (size_t)n_bins + 32) * sizeof(double);
    }
    __device__ WarpHist(double *smem, int nb) : all(smem), n_bins(nb) {
        const int n_warps = blockDim.x >> 5, warp = threadIdx.x >> 5;
        bins = smem + (size_t)warp * 2 * nb;
        stage = smem + (size_t)n_warps * 2 * nb + warp * 32;
    }
    __device__ void clear() {
        const int total = (blockDim.x >> 5) * 2 * n_bins;
        for (int i = threadIdx.x; i < total; i += blockDim.x) all[i] = 0.0;
        __syncthreads();
    }
    // Warp-collective: all 32 lanes must call (bin < 0 = nothing to add).
    __device__ __forceinline__ void add(int bin, double w) {
        const int lane = threadIdx.x & 31;
        __syncwarp();
        const unsigned peers = __match_any_sync(0xffffffffu, bin);
        stage[lane] = w;
        __syncwarp();
        if (bin >= 0 && lane == __ffs(peers) - 1) {
            double s = 0.0, s2 = 0.0;
            unsigned m = peers;
            while (m) {
                const int l = __ffs(m) - 1;
                m &= m - 1;
                const double x = stage[l];
                s += x;
                s2 = fma(x, x, s2);
            }
            bins[bin] += s;
            bins[n_bins + bin] += s2;
        }
        __syncwarp();
    }
    // Block-collective: sum the warps' copies in warp order into dst[2*n_bins].
    __device__ void flush(double *dst) {
        __syncthreads();
        const int n_warps = blockDim.x >> 5;
        for (int b = threadIdx.x; b < 2 * n_bins; b += blockDim.x) {
            double s = 0.0;
            for (int w = 0; w < n_warps; ++w) s += all[(size_t)w * 2 * n_bins + b];
            dst[b] = s;
        }
    }
};

// ---- large binnings (> PISAB_DET_MAX_BINS): exact fixed-point accumulation ---------------------------------
// Private copies of the bins no longer fit in shared memory, and floating-point atomics on shared bins are not
// reproducible (the order of the additions changes from run to run).  INTEGER addition is associative, so every
// weight is converted to a 128-bit two's-complement fixed-point number  w * 2^k = hi * 2^64 + lo  and added with two
// 64-bit integer atomics (the carry out of the low word is detected from the returned old value and added to the
// high word).  k is chosen from an upper bound of N * max|w| so that the sum cannot overflow; with N = 1e8 events a
// weight keeps all 53 bits of its mantissa as long as it is within 2^47 of the largest one.  The result is the
// EXACT sum rounded once: bit-reproducible on any grid, any number of GPUs' worth of event order, any run.
struct FixedAcc {
    unsigned long long lo, hi;
};
// scale = 2^(62 - ex) with 2^ex > n * bound  (so |sum| * scale < 2^62); bound == 0 (all weights zero) -> 1
__device__ __forceinline__ double fixed_scale(double bound, double n) {
    const double tot = bound * n;
    if (!(tot > 0.0)) return 1.0;
    int ex;
    frexp(tot, &ex);
    return ldexp(1.0, 62 - ex);
}
__device__ __forceinline__ void fixed_add(FixedAcc *acc, double v, double scale) {
    const double s = v * scale;               // exact: scale is a power of two
    const double fl = floor(s);
    const double rem = s - fl;                // [0, 1), exact
    const unsigned long long lo = (unsigned long long)(rem * 18446744073709551616.0);
    const long long hi = (long long)fl;
    const unsigned long long old = atomicAdd(&acc->lo, lo);
    const unsigned long long carry = (old + lo) < old ? 1ull : 0ull;
    const unsigned long long add_hi = (unsigned long long)hi + carry;
    if (add_hi) atomicAdd(&acc->hi, add_hi);
}
__device__ __forceinline__ double fixed_value(const FixedAcc &a, double scale) {
    return ((double)(long long)a.hi + (double)a.lo * 5.421010862427522e-20) / scale; // lo * 2^-64
}

// persistent grid for the histogramming kernels (fixed by the device -> reproducible sums)
int hist_grid(int64_t n);
// hist[b] = sum over blocks (in block order) of partials[block][b]; w2 likewise (nullable)
int hist_reduce_partials(const double *d_partials, int n_blocks, int n_bins, double *d_hist,
                         double *d_hist_w2, cudaStream_t s);

int hist_reduce_batch(const double *d_partials, int n_blocks, int n_bins, int n_containers, double *d_out,
                      cudaStream_t s);
// the same reduction + per-bin scales + container sum + mod_chi2 in one launch (last-block epilogue)
int hist_reduce_chi2(const double *d_partials, int n_blocks, int n_bins, int n_containers, const double *d_bin_scales,
                     const double *d_observed, double *d_out, double *d_total, double *d_chi2, unsigned *d_arrive,
                     cudaStream_t s);

} // namespace pisab
