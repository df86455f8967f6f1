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
    // Warp-collective: all 32 lanes must call (a bin outside [0, n_bins) = nothing to add).
    __device__ __forceinline__ void add(int bin, double w) {
        const int lane = threadIdx.x & 31;
        if ((unsigned)bin >= (unsigned)n_bins) bin = -1;
        // (no barrier here: every add ends with one, and clear() ends with a block barrier)
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
    // Two events per lane (the two-events-per-thread kernel): the same additions in the same order as add(bin0, w0)
    // followed by add(bin1, w1) -- every bin first receives its group sum of the first events, then of the second --
    // but the two matches, the staging and the two group sums overlap, and one warp barrier goes away; the histogram
    // tail was 9 % of that kernel's stall samples.  stage1: 32 more staging slots of this warp.
    __device__ __forceinline__ void add2(int bin0, double w0, int bin1, double w1, double *stage1) {
        const int lane = threadIdx.x & 31;
        if ((unsigned)bin0 >= (unsigned)n_bins) bin0 = -1;
        if ((unsigned)bin1 >= (unsigned)n_bins) bin1 = -1;
        const unsigned p0 = __match_any_sync(0xffffffffu, bin0), p1 = __match_any_sync(0xffffffffu, bin1);
        stage[lane] = w0;
        stage1[lane] = w1;
        __syncwarp();
        const bool lead0 = bin0 >= 0 && lane == __ffs(p0) - 1, lead1 = bin1 >= 0 && lane == __ffs(p1) - 1;
        double s = 0.0, s2 = 0.0, t = 0.0, t2 = 0.0;
        if (lead0) {
            unsigned m = p0;
            while (m) {
                const int l = __ffs(m) - 1;
                m &= m - 1;
                const double x = stage[l];
                s += x;
                s2 = fma(x, x, s2);
            }
        }
        if (lead1) {
            unsigned m = p1;
            while (m) {
                const int l = __ffs(m) - 1;
                m &= m - 1;
                const double x = stage1[l];
                t += x;
                t2 = fma(x, x, t2);
            }
        }
        if (lead0) {
            bins[bin0] += s;
            bins[n_bins + bin0] += s2;
        }
        __syncwarp();
        if (lead1) {
            bins[bin1] += t;
            bins[n_bins + bin1] += t2;
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
// weight is converted to a 128-bit two's-complement fixed-point number  w * 2^k = hi * 2^64 + lo  and added with
// integer atomics.  k is chosen from an upper bound of N * max|w| so that the sum cannot overflow; with N = 1e8 events
// a weight keeps all 53 bits of its mantissa as long as it is within 2^47 of the largest one.  The result is the
// EXACT sum rounded once: bit-reproducible on any grid, any number of GPUs' worth of event order, any run.
// The accumulator is kept in CARRY-SAVE form -- the two 32-bit halves of `lo` are summed in 64-bit words of their own
// (room for 2^32 additions) -- so that an addition is three fire-and-forget reductions (RED) instead of an atomic whose
// returned value decides the carry: the issuing warp never waits for the L2 round trip.
struct FixedAcc {
    unsigned long long lo0, lo1, hi, pad; // value = hi * 2^64 + lo1 * 2^32 + lo0 (mod 2^128), 32-byte entries
};
// scale = 2^(62 - ex) with 2^ex > n * bound  (so |sum| * scale < 2^62); bound == 0 (all weights zero) -> 1
__device__ __forceinline__ double fixed_scale(double bound, double n) {
    const double tot = bound * n;
    if (!(tot > 0.0)) return 1.0;
    int ex;
    frexp(tot, &ex);
    return ldexp(1.0, 62 - ex);
}
struct Fixed128 { // one weight (or a sum of weights) as a 128-bit two's-complement fixed-point number
    unsigned long long lo;
    long long hi;
};
__device__ __forceinline__ Fixed128 to_fixed(double v, double scale) {
    const double s = v * scale;               // exact: scale is a power of two
    const double fl = floor(s);
    const double rem = s - fl;                // [0, 1), exact
    Fixed128 r;
    r.lo = (unsigned long long)(rem * 18446744073709551616.0);
    r.hi = (long long)fl;
    return r;
}
__device__ __forceinline__ void fixed_sum(Fixed128 &a, const Fixed128 &b) { // a += b (integer, associative)
    const unsigned long long lo = a.lo + b.lo;
    a.hi += b.hi + (lo < a.lo ? 1 : 0);
    a.lo = lo;
}
__device__ __forceinline__ void fixed_add(FixedAcc *acc, const Fixed128 &x) {
    atomicAdd(&acc->lo0, x.lo & 0xffffffffull); // results unused: RED, not ATOM
    atomicAdd(&acc->lo1, x.lo >> 32);
    if (x.hi) atomicAdd(&acc->hi, (unsigned long long)x.hi);
}
__device__ __forceinline__ void fixed_add(FixedAcc *acc, double v, double scale) { fixed_add(acc, to_fixed(v, scale)); }
__device__ __forceinline__ Fixed128 warp_fixed_reduce(Fixed128 v) { // sum over the 32 lanes (xor tree), in every lane
    for (int off = 16; off > 0; off >>= 1) {
        Fixed128 o;
        o.lo = __shfl_xor_sync(0xffffffffu, v.lo, off);
        o.hi = __shfl_xor_sync(0xffffffffu, v.hi, off);
        fixed_sum(v, o);
    }
    return v;
}

// Warp-collective form for events that arrive grouped by bin (the engine orders the events of a large binning by
// bin inside every class of crossed shells).  All 32 lanes in one bin -- the usual case then: the fixed-point weights
// are summed by a shuffle tree (integer addition: still exact and order-independent) and lane 0 issues ONE set of
// reductions per plane instead of one per event.  Otherwise lanes that share a bin are grouped with __match_any_sync
// and the lowest lane of each group sums its group through the warp's staging words.
// stage: this warp's [4][32] unsigned long long.  All 32 lanes must call.
__device__ __forceinline__ void warp_fixed_add(unsigned long long *stage, FixedAcc *acc, int n_bins, int bin, double w,
                                               double sc1, double sc2) {
    const int lane = threadIdx.x & 31;
    const bool ok = (unsigned)bin < (unsigned)n_bins;
    const int key = ok ? bin : -1;
    const Fixed128 a = to_fixed(w, sc1), b = to_fixed(w * w, sc2);
    const int key0 = __shfl_sync(0xffffffffu, key, 0);
    if (__all_sync(0xffffffffu, key == key0)) {
        if (key0 < 0) return;
        const Fixed128 s1 = warp_fixed_reduce(a), s2 = warp_fixed_reduce(b);
        if (lane == 0) {
            fixed_add(acc + key0, s1);
            fixed_add(acc + n_bins + key0, s2);
        }
        return;
    }
    __syncwarp();
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    stage[lane] = a.lo;
    stage[32 + lane] = (unsigned long long)a.hi;
    stage[64 + lane] = b.lo;
    stage[96 + lane] = (unsigned long long)b.hi;
    __syncwarp();
    if (ok && lane == __ffs(peers) - 1) {
        Fixed128 s1 = a, s2 = b;
        unsigned m = peers & (peers - 1); // the other lanes of the group
        while (m) {
            const int l = __ffs(m) - 1;
            m &= m - 1;
            fixed_sum(s1, Fixed128{stage[l], (long long)stage[32 + l]});
            fixed_sum(s2, Fixed128{stage[64 + l], (long long)stage[96 + l]});
        }
        fixed_add(acc + bin, s1);
        fixed_add(acc + n_bins + bin, s2);
    }
    __syncwarp();
}
// carry-save words -> hi * 2^64 + lo, then to double
__device__ __forceinline__ double fixed_value(const FixedAcc &a, double scale) {
    const unsigned long long mid = a.lo1 << 32;
    const unsigned long long lo = a.lo0 + mid;
    const unsigned long long hi = a.hi + (a.lo1 >> 32) + (lo < mid ? 1ull : 0ull);
    return ((double)(long long)hi + (double)lo * 5.421010862427522e-20) / scale; // lo * 2^-64
}

// ---- fit-loop epilogue (reduce the per-block partial histograms, per-bin systematics scales, container sum, mod_chi2)
// Shared by hist_reduce_chi2_kernel (a launch of its own) and the template kernels, whose last-arriving blocks run it
// so that one hypothesis of a fit is ONE launch.
struct FusedEpi {
    const double *bin_scales, *observed; // optional [n_containers][n_bins] / [n_bins]
    double *out;                         // [n_containers][2][n_bins]; nullptr: no epilogue in this launch
    double *total, *chi2;                // optional [2][n_bins] / one double
    unsigned *arrive;                    // zero-initialised words: [0] containers done, [1 + c] blocks of container c done
};

// value v = (plane, bin) of container c: sum of its n_blocks partials -- lane l adds blocks l, l+32, ... in ascending
// order, the 32 lane sums are combined by a fixed xor tree (the order depends on n_blocks only) -- then the optional
// discr_sys.hypersurfaces scale (hypersurfaces.py:219-248: weights -> clip(s w, 0), errors -> s errors).
__device__ __forceinline__ double apply_bin_scale(double r, int v, int n_bins, const double *scales_c) {
    if (!scales_c) return r;
    const double sc = scales_c[v < n_bins ? v : v - n_bins];
    return v < n_bins ? fmax(r * sc, 0.0) : r * sc * sc;
}
// one warp per value (the stand-alone reduction kernels: thousands of warps, every lane one strided load)
__device__ __forceinline__ void reduce_container_value(const double *partials_c, int n_blocks, int n_bins, int v,
                                                       const double *scales_c, double *out_c) {
    const int lane = threadIdx.x & 31, n_values = 2 * n_bins;
    double s = 0.0;
    for (int k = lane; k < n_blocks; k += 32) s += __ldcg(partials_c + (size_t)k * n_values + v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out_c[v] = apply_bin_scale(s, v, n_bins, scales_c);
}
// one THREAD per value, block-collective (the last block of a container inside the template kernel): consecutive
// threads read consecutive values of one partial histogram -- coalesced, where the warp-per-value form makes one block
// fetch 32 sectors per load instruction (10 us for 24 partials, 1 ms for 1579).  The same sums in the same order:
// 32 register accumulators play the lanes, and since a + b == b + a the xor tree folds to t[l] += t[l + o].
__device__ __forceinline__ void reduce_container_block(const double *partials_c, int n_blocks, int n_bins,
                                                       const double *scales_c, double *out_c) {
    const int n_values = 2 * n_bins;
    for (int v = threadIdx.x; v < n_values; v += blockDim.x) {
        double acc[32];
#pragma unroll
        for (int l = 0; l < 32; ++l) acc[l] = 0.0;
        for (int k0 = 0; k0 < n_blocks; k0 += 32) {
#pragma unroll
            for (int h = 0; h < 32; h += 16) { // 16 loads in flight (32 would spill next to the 32 accumulators)
                double x[16];
#pragma unroll
                for (int l = 0; l < 16; ++l)
                    x[l] = k0 + h + l < n_blocks ? __ldcg(partials_c + (size_t)(k0 + h + l) * n_values + v) : 0.0;
#pragma unroll
                for (int l = 0; l < 16; ++l)
                    if (k0 + h + l < n_blocks) acc[h + l] += x[l];
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int l = 0; l < o; ++l) acc[l] += acc[l + o];
        }
        out_c[v] = apply_bin_scale(acc[0], v, n_bins, scales_c);
    }
}

// MapSet sum over the containers (fixed order) and mod_chi2 (stats.py:651-695).  Block-collective; scratch: at least
// 2 * n_bins + blockDim.x doubles of shared memory.  One thread per (plane, bin) issues the loads of all containers
// together: this is the serial tail of every template.
__device__ __forceinline__ void template_total_chi2(const double *out, int n_containers, int n_bins, const double *observed,
                                                    double *total, double *chi2, double *scratch) {
    double *s_tot = scratch, *s_acc = scratch + 2 * n_bins;
    for (int v = threadIdx.x; v < 2 * n_bins; v += blockDim.x) {
        double x[PISAB_MAX_BATCH];
#pragma unroll
        for (int c = 0; c < PISAB_MAX_BATCH; ++c) x[c] = c < n_containers ? __ldcg(out + (size_t)c * 2 * n_bins + v) : 0.0;
        double t = 0.0;
#pragma unroll
        for (int c = 0; c < PISAB_MAX_BATCH; ++c) t += x[c]; // (absent containers add +0.0)
        s_tot[v] = t;
        if (total) total[v] = t;
    }
    __syncthreads();
    double acc = 0.0;
    if (observed) {
        for (int b = threadIdx.x; b < n_bins; b += blockDim.x) {
            const double e = fmax(s_tot[b], 1e-10), sig2 = s_tot[n_bins + b];
            const double d = observed[b] - e;
            acc += d * d / (sig2 + e);
        }
    }
    s_acc[threadIdx.x] = acc;
    __syncthreads();
    for (int off = blockDim.x >> 1; off > 0; off >>= 1) {
        if (threadIdx.x < off) s_acc[threadIdx.x] += s_acc[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0 && chi2) chi2[0] = s_acc[0];
}

// Called by every block of a template kernel after it has written its partial histogram of container ci (rank of
// `ranks`).  The last block of a container reduces that container; the last container to finish sums the containers
// and evaluates the chi2.  scratch: the block's (now idle) dynamic shared memory.
__device__ __forceinline__ void fused_epilogue(const FusedEpi &E, const double *partials, int ci, int ranks, int n_containers,
                                               int n_bins, double *scratch) {
    __shared__ int s_role;
    __threadfence(); // this block's partial histogram is visible device-wide before it is counted
    __syncthreads();
    if (threadIdx.x == 0) s_role = atomicAdd(E.arrive + 1 + ci, 1u) == (unsigned)ranks - 1 ? 1 : 0;
    __syncthreads();
    if (!s_role) return;
    __threadfence();
    const int n_values = 2 * n_bins;
    reduce_container_block(partials + (size_t)ci * ranks * n_values, ranks, n_bins,
                           E.bin_scales ? E.bin_scales + (size_t)ci * n_bins : nullptr, E.out + (size_t)ci * n_values);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        E.arrive[1 + ci] = 0; // the next launch on this stream starts from zero
        s_role = atomicAdd(E.arrive, 1u) == (unsigned)n_containers - 1 ? 2 : 0;
    }
    __syncthreads();
    if (s_role != 2) return;
    if (threadIdx.x == 0) E.arrive[0] = 0;
    __threadfence();
    if (E.total || E.chi2) template_total_chi2(E.out, n_containers, n_bins, E.observed, E.total, E.chi2, scratch);
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
// zero-initialised arrival words of the epilogue for (current device, stream); nullptr when none can be had
unsigned *epilogue_counter(cudaStream_t s);

} // namespace pisab
