// hist.cu -- bin indexing, weighted histogram accumulation, lookup, small element-wise ops.
#include <math.h>

#include "hist_device.cuh"

namespace pisab {

// ---------------------------------------------------------------------------------------------
// bin index (translation.py:417-455 rule; hist.py:93-113 for irregular dims)
// ---------------------------------------------------------------------------------------------
constexpr int kSlotsMaxBins = 512; // above: match-based warp-private scheme

struct BinningTable {
    int n_dims;
    int kind[PISAB_MAX_DIMS];
    int n_bins[PISAB_MAX_DIMS];
    double lo[PISAB_MAX_DIMS], hi[PISAB_MAX_DIMS], norm[PISAB_MAX_DIMS];
    const double *edges[PISAB_MAX_DIMS];
};

template <typename IO>
struct CoordPtrs {
    const IO *p[PISAB_MAX_DIMS];
};

// np.searchsorted(edges, x, side='right') - 1, with x == last edge folded into the last bin
__device__ __forceinline__ int digitize_right(const double *__restrict__ edges, int n_edges, double v) {
    if (v != v) return n_edges - 1; // nan sorts last: searchsorted returns n_edges
    int lo = 0, hi = n_edges;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (edges[mid] <= v) lo = mid + 1;
        else hi = mid;
    }
    int idx = lo - 1;
    if (v == edges[n_edges - 1]) idx -= 1;
    return idx;
}

// One dimension of one event: bin id and in-range flag (shared by the generic and unrolled kernels)
template <typename IO>
__device__ __forceinline__ int index_1d(const BinningTable &B, int d, double x, bool &ok) {
    // arithmetic is double in both storage modes (numba promotes float32 op float64; fast_histogram
    // converts to double)
    if (B.kind[d] == PISAB_DIM_EDGES) {
        const int id = digitize_right(B.edges[d], B.n_bins[d] + 1, x);
        ok = ok && id >= 0 && id < B.n_bins[d];
        return id;
    }
    // Container.translate: np.log on the FTYPE array (container.py:845-850)
    if (B.kind[d] == PISAB_DIM_LOG) x = sizeof(IO) == 4 ? (double)logf((float)x) : log(x);
    const bool in = x >= B.lo[d] && x < B.hi[d];
    // (int)((x - lo) * norm): no FMA contraction
    int id = in ? (int)__dmul_rn(__dsub_rn(x, B.lo[d]), B.norm[d]) : 0;
    // x < hi whose product rounds up to n: out-of-bounds access in the reference, folded into the last
    // bin here and in the oracle
    if (id >= B.n_bins[d]) id = B.n_bins[d] - 1;
    ok = ok && in;
    return id;
}

// NDIMS is a compile-time constant (1..PISAB_MAX_DIMS) so the dimension loop unrolls, and every thread
// handles two events per iteration so that 2 NDIMS loads are in flight.  HBM-bound: 8 NDIMS B in + 4 B out.
constexpr int kSmemEdges = 128; // per dimension; longer edge lists are searched in global memory

template <typename IO, int NDIMS>
__global__ void __launch_bounds__(256)
hist_index_kernel(const __grid_constant__ BinningTable Bg, const __grid_constant__ CoordPtrs<IO> C,
                  int64_t n, int32_t *__restrict__ out) {
    // explicit edge lists (irregular dimensions) are searched per event: stage them in shared memory
    __shared__ double s_edges[NDIMS][kSmemEdges];
    BinningTable B = Bg;
#pragma unroll
    for (int d = 0; d < NDIMS; ++d) {
        if (B.kind[d] == PISAB_DIM_EDGES && B.n_bins[d] + 1 <= kSmemEdges) {
            for (int k = threadIdx.x; k <= B.n_bins[d]; k += blockDim.x) s_edges[d][k] = __ldg(Bg.edges[d] + k);
            B.edges[d] = s_edges[d];
        }
    }
    __syncthreads();
    // LOG dimensions arrive with their RAW domain: its logarithm is taken here, by the same function (and in the
    // same precision) as the samples' below, so that a sample equal to an edge lands on the edge -- like the
    // reference, whose regularised domain and samples both go through np.log of FTYPE values (hist.py:118-120,
    // container.py:845-850).  Host-side logs can differ from the device's by an ulp.
#pragma unroll
    for (int d = 0; d < NDIMS; ++d) {
        if (B.kind[d] == PISAB_DIM_LOG) {
            B.lo[d] = sizeof(IO) == 4 ? (double)logf((float)B.lo[d]) : log(B.lo[d]);
            B.hi[d] = sizeof(IO) == 4 ? (double)logf((float)B.hi[d]) : log(B.hi[d]);
            B.norm[d] = (double)B.n_bins[d] / (B.hi[d] - B.lo[d]);
        }
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < n; i += 2 * stride) {
        double xa[NDIMS], xb[NDIMS];
#pragma unroll
        for (int d = 0; d < NDIMS; ++d) {
            xa[d] = (double)__ldg(C.p[d] + i);
            xb[d] = (double)__ldg(C.p[d] + i + stride);
        }
        int fa = 0, fb = 0;
        bool oka = true, okb = true;
#pragma unroll
        for (int d = 0; d < NDIMS; ++d) {
            fa = fa * B.n_bins[d] + index_1d<IO>(B, d, xa[d], oka);
            fb = fb * B.n_bins[d] + index_1d<IO>(B, d, xb[d], okb);
        }
        out[i] = oka ? fa : -1;
        out[i + stride] = okb ? fb : -1;
    }
    for (; i < n; i += stride) {
        int flat = 0;
        bool ok = true;
#pragma unroll
        for (int d = 0; d < NDIMS; ++d) flat = flat * B.n_bins[d] + index_1d<IO>(B, d, (double)__ldg(C.p[d] + i), ok);
        out[i] = ok ? flat : -1;
    }
}

// ---------------------------------------------------------------------------------------------
// accumulation
// ---------------------------------------------------------------------------------------------
template <typename IO>
__global__ void __launch_bounds__(kHistBlock)
hist_accumulate_kernel(const int32_t *__restrict__ index, const IO *__restrict__ weights, int64_t n,
                       int n_bins, double *__restrict__ partials) {
    extern __shared__ double s_hist[];
    WarpHist wh(s_hist, n_bins);
    wh.clear();
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t warp_first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x - lane;
    // 4 warp-steps per iteration so that 8 loads are in flight per lane
    int64_t base = warp_first;
    for (; base + 3 * stride < n; base += 4 * stride) {
        int b[4];
        double w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = base + u * stride + lane;
            const bool ok = i < n;
            b[u] = ok ? __ldg(index + i) : -1;
            w[u] = ok ? (weights ? (double)__ldg(weights + i) : 1.0) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) wh.add(b[u], w[u]);
    }
    for (; base < n; base += stride) {
        const int64_t i = base + lane;
        const bool ok = i < n;
        const int b = ok ? __ldg(index + i) : -1;
        const double w = ok ? (weights ? (double)__ldg(weights + i) : 1.0) : 0.0;
        wh.add(b, w);
    }
    wh.flush(partials + (size_t)blockIdx.x * 2 * n_bins);
}

// ---- small binnings (the analysis binnings: 8x8x2 = 128 bins): lane-replicated private bins ----
// Every WARP owns R copies ("slots") of the bins, slot = lane % R, each entry a (sum w, sum w^2) pair:
//     s_bins[warp][bin][slot]  (16 bytes)  ->  R * n_bins * 16 bytes per warp.
// Lanes with different slots can never collide, so the 32 lanes are served in 32/R passes of R lanes;
// inside a pass each lane does a plain 128-bit read-modify-write per event (LDS.128, DADD, DFMA,
// STS.128 -- bank-conflict free: bank = 4 * slot) for all U events it has in registers, and only the
// pass boundary needs a __syncwarp().  No match, no staging, no atomics: ~50 instructions per 32 events
// instead of ~100 for the match-based scheme below (__match_any_sync alone costs more than a whole
// step here; variants with match or xor-shuffle de-duplication over a slot column, and with paired
// read-modify-writes, were measured slower: profiles/r01_hist_variants.txt), and the summation order
// is fixed by the launch geometry alone (bit-reproducible).  A TMA-fed variant (ring of cp.async.bulk
// tiles + mbarriers, one block per SM) was also measured: 0.49 ms vs 0.44 ms for this kernel at 1e8
// events -- once the data delivery is solved the limit is the 2 warps per scheduler that 16 KB of private
// bins per warp allow, so it was not kept.
template <typename IO, int R, int U>
__global__ void __launch_bounds__(kHistBlock)
hist_accumulate_slots_kernel(const int32_t *__restrict__ index, const IO *__restrict__ weights, int64_t n,
                             int n_bins, double *__restrict__ partials) {
    extern __shared__ __align__(16) double2 s_slots[]; // [warps][n_bins][R]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    double2 *mine = s_slots + (size_t)warp * n_bins * R + (lane % R);
    for (int i = threadIdx.x; i < n_warps * n_bins * R; i += blockDim.x) s_slots[i] = make_double2(0.0, 0.0);
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t warp_first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x - lane;
    const int group = lane / R;
    auto load_tile = [&](int64_t base, int (&b)[U], double (&w)[U]) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * stride + lane;
            const bool ok = i < n;
            const int bb = ok ? __ldg(index + i) : -1;
            b[u] = (unsigned)bb < (unsigned)n_bins ? bb : -1;
            w[u] = ok ? (weights ? (double)__ldg(weights + i) : 1.0) : 0.0;
        }
    };
    auto bin_tile = [&](const int (&b)[U], const double (&w)[U]) {
#pragma unroll
        for (int p = 0; p < 32 / R; ++p) {
            if (group == p) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (b[u] >= 0) {
                        double2 *slot = mine + (size_t)b[u] * R;
                        double2 v = *slot;
                        v.x += w[u];
                        v.y = fma(w[u], w[u], v.y);
                        *slot = v;
                    }
                }
            }
            __syncwarp();
        }
    };
    // double buffered: the loads of tile k+1 are in flight while tile k is binned (the kernel is bound by
    // DRAM latency x bytes in flight: 12 warps/SM x U x 384 B)
    const int64_t tile = (int64_t)U * stride;
    int b0[U], b1[U];
    double w0[U], w1[U];
    load_tile(warp_first, b0, w0);
    for (int64_t base = warp_first; base < n; base += 2 * tile) { // warp-uniform trip count
        load_tile(base + tile, b1, w1);
        bin_tile(b0, w0);
        load_tile(base + 2 * tile, b0, w0);
        bin_tile(b1, w1);
    }
    __syncthreads();
    // block partial: slots in ascending order, then warps in ascending order (fixed)
    double *dst = partials + (size_t)blockIdx.x * 2 * n_bins;
    for (int bin = threadIdx.x; bin < n_bins; bin += blockDim.x) {
        double sw = 0.0, sw2 = 0.0;
        for (int wp = 0; wp < n_warps; ++wp) {
            const double2 *src = s_slots + ((size_t)wp * n_bins + bin) * R;
            for (int r = 0; r < R; ++r) { sw += src[r].x; sw2 += src[r].y; }
        }
        dst[bin] = sw;
        dst[n_bins + bin] = sw2;
    }
}

// ---- planned histogram: bin-sorted tiles prepared once, no private bins at all -------------------------------
// The bin index of an event never changes during a fit (hist.setup_function computes it once), so the expensive
// part of histogramming -- routing 32 weights of a warp to 32 arbitrary bins without collisions -- can be done ONCE:
// pisab_hist_plan_build cuts the events into tiles of 2048, and stores for every tile the permutation that lists its
// events grouped by bin (uint16 positions inside the tile, stable) plus the n_bins + 1 group offsets.  The per-template
// kernel then stages a tile's weights (coalesced cp.async, 8 B/event) and its 2 B/event permutation in shared
// memory, and THREAD b walks group b: two shared-memory loads, one DADD and one DFMA per event into REGISTER
// accumulators that thread b keeps for the whole launch.  No read-modify-write on bins, no atomics, no replication
// (the 16 KB of replicated bins per warp is what capped the slot kernel at 12 warps per SM), 10.1 B/event of DRAM
// traffic instead of 12, and the summation order is fixed by the plan: bit-reproducible.
constexpr int kPlanTile = 2048;
constexpr int kPlanMaxBins = PISAB_DET_MAX_BINS; // 1024: up to four bins per thread
struct PlanLayout {
    int64_t n_tiles;
    int off_stride;     // uint16 entries per tile in the offsets table (n_bins + 1 rounded up to 8: 16-byte rows)
    size_t off_bytes, perm_bytes;
    __host__ __device__ PlanLayout(int64_t n, int n_bins) {
        n_tiles = (n + kPlanTile - 1) / kPlanTile;
        off_stride = (n_bins + 1 + 7) / 8 * 8;
        off_bytes = ((size_t)n_tiles * off_stride * 2 + 255) / 256 * 256;
        perm_bytes = (size_t)n_tiles * kPlanTile * 2;
    }
    __host__ __device__ size_t bytes() const { return off_bytes + perm_bytes; }
};

// one block per tile; thread b counts, then places, the events of bins b, b + blockDim, ... in tile order (stable)
__global__ void __launch_bounds__(256)
hist_plan_kernel(const int32_t *__restrict__ index, int64_t n, int n_bins, uint16_t *__restrict__ offsets, int off_stride,
                 uint16_t *__restrict__ perm) {
    __shared__ int16_t s_bin[kPlanTile];
    __shared__ int s_count[kPlanMaxBins + 1];
    const int64_t tile = blockIdx.x, base = tile * kPlanTile;
    for (int i = threadIdx.x; i < kPlanTile; i += blockDim.x) {
        const int64_t g = base + i;
        const int b = g < n ? __ldg(index + g) : -1;
        s_bin[i] = (int16_t)((unsigned)b < (unsigned)n_bins ? b : -1);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < n_bins; b += blockDim.x) {
        int cnt = 0;
        for (int i = 0; i < kPlanTile; ++i) cnt += (s_bin[i] == b);
        s_count[b] = cnt;
    }
    __syncthreads();
    if (threadIdx.x < 32) { // exclusive scan by one warp: lane l owns a contiguous run of bins
        const int per = (n_bins + 31) / 32, lo = threadIdx.x * per, hi = min(lo + per, n_bins);
        int mine = 0;
        for (int k = lo; k < hi; ++k) mine += s_count[k];
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)threadIdx.x >= o) incl += v;
        }
        int acc = incl - mine;
        for (int k = lo; k < hi; ++k) { const int c = s_count[k]; s_count[k] = acc; acc += c; }
        if (threadIdx.x == 31) s_count[n_bins] = incl;
    }
    __syncthreads();
    uint16_t *my_perm = perm + tile * kPlanTile;
    for (int b = threadIdx.x; b < n_bins; b += blockDim.x) {
        int pos = s_count[b];
        for (int i = 0; i < kPlanTile; ++i)
            if (s_bin[i] == b) my_perm[pos++] = (uint16_t)i;
    }
    for (int k = threadIdx.x; k <= n_bins; k += blockDim.x) offsets[tile * off_stride + k] = (uint16_t)s_count[k];
    // (entries of perm beyond offsets[n_bins] are never read)
}

template <typename IO, int THREADS, int BPT> // thread t owns bins t, t + THREADS, ... (BPT of them)
__global__ void __launch_bounds__(THREADS)
hist_planned_kernel(const uint16_t *__restrict__ offsets, int off_stride, const uint16_t *__restrict__ perm,
                    const IO *__restrict__ weights, int64_t n, int n_bins, double *__restrict__ partials) {
    extern __shared__ __align__(16) unsigned char s_plan[];
    // two stages of { weights[kPlanTile] (IO), perm[kPlanTile] (u16), offsets[off_stride] (u16) }
    const size_t w_bytes = (size_t)kPlanTile * sizeof(IO), p_bytes = (size_t)kPlanTile * 2;
    const size_t stage_bytes = w_bytes + p_bytes + (size_t)off_stride * 2;
    const int64_t n_tiles = (n + kPlanTile - 1) / kPlanTile;
    const int tid = threadIdx.x;
    auto stage_ptr = [&](int st) { return s_plan + (size_t)st * stage_bytes; };
    auto issue = [&](int64_t tile, int st) {
        unsigned char *dst = stage_ptr(st);
        const int64_t base = tile * kPlanTile;
        const int64_t left = n - base; // events in this tile
        const IO *src_w = weights + base;
        if (left >= kPlanTile) {
            for (int c = tid; c < (int)(w_bytes / 16); c += THREADS) {
                const unsigned sm = (unsigned)__cvta_generic_to_shared(dst + (size_t)c * 16);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sm), "l"((const char *)src_w + (size_t)c * 16) : "memory");
            }
        } else { // last, partial tile: guarded element loads
            IO *dw = reinterpret_cast<IO *>(dst);
            for (int i = tid; i < kPlanTile; i += THREADS) dw[i] = i < left ? __ldg(src_w + i) : (IO)0;
        }
        const uint16_t *src_p = perm + tile * kPlanTile;
        for (int c = tid; c < (int)(p_bytes / 16); c += THREADS) {
            const unsigned sm = (unsigned)__cvta_generic_to_shared(dst + w_bytes + (size_t)c * 16);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sm), "l"((const char *)src_p + (size_t)c * 16) : "memory");
        }
        const uint16_t *src_o = offsets + tile * off_stride;
        for (int c = tid; c < off_stride / 8; c += THREADS) {
            const unsigned sm = (unsigned)__cvta_generic_to_shared(dst + w_bytes + p_bytes + (size_t)c * 16);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sm), "l"((const char *)src_o + (size_t)c * 16) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    double sa[BPT], sb[BPT], qa[BPT], qb[BPT]; // two interleaved chains per sum: order fixed by (plan, grid)
#pragma unroll
    for (int j = 0; j < BPT; ++j) sa[j] = sb[j] = qa[j] = qb[j] = 0.0;
    int64_t tile = blockIdx.x;
    int st = 0;
    if (tile < n_tiles) issue(tile, 0);
    for (; tile < n_tiles; tile += gridDim.x, st ^= 1) {
        const int64_t next = tile + gridDim.x;
        if (next < n_tiles) {
            issue(next, st ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const unsigned char *base = stage_ptr(st);
        const IO *w = reinterpret_cast<const IO *>(base);
        const uint16_t *p = reinterpret_cast<const uint16_t *>(base + w_bytes);
        const uint16_t *o = reinterpret_cast<const uint16_t *>(base + w_bytes + p_bytes);
#pragma unroll
        for (int j = 0; j < BPT; ++j) {
            const int b = tid + j * THREADS;
            if (b < n_bins) {
                int k = o[b];
                const int end = o[b + 1];
                for (; k + 1 < end; k += 2) {
                    const double x = (double)w[p[k]], y = (double)w[p[k + 1]];
                    sa[j] += x; qa[j] = fma(x, x, qa[j]);
                    sb[j] += y; qb[j] = fma(y, y, qb[j]);
                }
                if (k < end) {
                    const double x = (double)w[p[k]];
                    sa[j] += x; qa[j] = fma(x, x, qa[j]);
                }
            }
        }
        __syncthreads(); // the stage is refilled two iterations from now, by the issue() of the next iteration
    }
    double *dst = partials + (size_t)blockIdx.x * 2 * n_bins;
#pragma unroll
    for (int j = 0; j < BPT; ++j) {
        const int b = tid + j * THREADS;
        if (b < n_bins) {
            dst[b] = sa[j] + sb[j];
            dst[n_bins + b] = qa[j] + qb[j];
        }
    }
}

// large binnings (> PISAB_DET_MAX_BINS): exact 128-bit fixed-point accumulation with integer atomics (hist_device.cuh)
// -- bit-reproducible whatever the order of the additions.  Three small kernels: max |w| (the scale must make the
// sum fit), accumulate, convert.
template <typename IO>
__global__ void __launch_bounds__(256)
absmax_kernel(const IO *__restrict__ weights, int64_t n, unsigned long long *__restrict__ bound_bits) {
    __shared__ double s[256];
    double m = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) m = fmax(m, fabs((double)__ldg(weights + i)));
    s[threadIdx.x] = m;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) s[threadIdx.x] = fmax(s[threadIdx.x], s[threadIdx.x + off]);
        __syncthreads();
    }
    // non-negative doubles order like their bit patterns; max is associative: deterministic
    if (threadIdx.x == 0) atomicMax(bound_bits, (unsigned long long)__double_as_longlong(s[0]));
}

template <typename IO>
__global__ void __launch_bounds__(256)
hist_accumulate_fixed_kernel(const int32_t *__restrict__ index, const IO *__restrict__ weights, int64_t n, int n_bins,
                             const unsigned long long *__restrict__ bound_bits, FixedAcc *__restrict__ acc, bool want_w2) {
    const double bound = weights ? __longlong_as_double((long long)*bound_bits) : 1.0;
    const double sc1 = fixed_scale(bound, (double)n), sc2 = fixed_scale(bound * bound, (double)n);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int b = __ldg(index + i);
        if (b < 0 || b >= n_bins) continue;
        const double w = weights ? (double)__ldg(weights + i) : 1.0;
        fixed_add(acc + b, w, sc1);
        if (want_w2) fixed_add(acc + n_bins + b, w * w, sc2);
    }
}

// Large binnings with a SORTED PLAN (setup: perm = stable order of the events by bin, sorted_index = index[perm]).
// Every warp owns a contiguous chunk of kSortedChunk plan entries; while the bin stays the same -- almost always: a bin
// of a 3200-bin histogram over 1e8 events holds 3e4 consecutive entries -- each lane adds its fixed-point weights in
// registers (integer addition: exact, associative), and only when the bin changes or the chunk ends are the 32 lanes
// combined (xor tree of 128-bit integer adds) and ONE atomic pair per plane issued.  ~1e5 atomics per 1e8 events
// instead of 4e8 (hist_accumulate_fixed_kernel: atomic-bound at 4.8 ms); what remains is 8 B/event of plan read
// coalesced plus one 32-byte sector per gathered weight.
constexpr int kSortedChunk = 2048;
template <typename IO>
__global__ void __launch_bounds__(256)
hist_accumulate_sorted_kernel(const int32_t *__restrict__ perm, const int32_t *__restrict__ sorted_index,
                              const IO *__restrict__ weights, int64_t n, int n_bins,
                              const unsigned long long *__restrict__ bound_bits, FixedAcc *__restrict__ acc) {
    __shared__ unsigned long long s_stage[8][128];
    const double bound = __longlong_as_double((long long)*bound_bits);
    const double sc1 = fixed_scale(bound, (double)n), sc2 = fixed_scale(bound * bound, (double)n);
    const int lane = threadIdx.x & 31;
    const int64_t n_chunks = (n + kSortedChunk - 1) / kSortedChunk;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t chunk = warp0; chunk < n_chunks; chunk += n_warps) {
        const int64_t end = (chunk + 1) * kSortedChunk < n ? (chunk + 1) * kSortedChunk : n;
        int run_bin = -1; // the bin the lanes' register sums belong to (-1: none)
        Fixed128 r1 = {0, 0}, r2 = {0, 0};
        auto flush = [&]() {
            if (run_bin >= 0) {
                const Fixed128 t1 = warp_fixed_reduce(r1), t2 = warp_fixed_reduce(r2);
                if (lane == 0) {
                    fixed_add(acc + run_bin, t1);
                    fixed_add(acc + n_bins + run_bin, t2);
                }
            }
            r1 = Fixed128{0, 0};
            r2 = Fixed128{0, 0};
            run_bin = -1;
        };
        for (int64_t base = chunk * kSortedChunk; base < end; base += 32) { // warp-uniform
            const int64_t i = base + lane;
            int b = -1;
            double w = 0.0;
            if (i < end) {
                b = __ldg(sorted_index + i);
                if ((unsigned)b < (unsigned)n_bins) w = (double)__ldg(weights + __ldg(perm + i));
                else b = -1;
            }
            const int b0 = __shfl_sync(0xffffffffu, b, 0);
            if (__all_sync(0xffffffffu, b == b0)) {
                if (b0 != run_bin) { flush(); run_bin = b0; }
                if (b0 >= 0) {
                    fixed_sum(r1, to_fixed(w, sc1));
                    fixed_sum(r2, to_fixed(w * w, sc2));
                }
            } else { // a bin boundary (or the ragged end) inside these 32 entries: one atomic pair per group
                flush();
                warp_fixed_add(s_stage[threadIdx.x >> 5], acc, n_bins, b, w, sc1, sc2);
            }
        }
        flush();
    }
}

__global__ void __launch_bounds__(256)
fixed_finish_kernel(const FixedAcc *__restrict__ acc, int n_bins, const unsigned long long *__restrict__ bound_bits,
                    bool has_weights, double n, double *__restrict__ hist, double *__restrict__ hist_w2) {
    const double bound = has_weights ? __longlong_as_double((long long)*bound_bits) : 1.0;
    const double sc1 = fixed_scale(bound, n), sc2 = fixed_scale(bound * bound, n);
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_bins) return;
    hist[b] = fixed_value(acc[b], sc1);
    if (hist_w2) hist_w2[b] = fixed_value(acc[n_bins + b], sc2);
}

// One warp per output value: lane l sums the partials of blocks l, l+32, ... in ascending order, the
// 32 lane sums are combined by a fixed xor tree -- the summation order depends on (n_blocks) only,
// so the result is bit-reproducible, and the 2*n_bins chains run in parallel instead of one
// thread walking all blocks (32 us -> ~3 us for 296 blocks x 256 values).
__global__ void __launch_bounds__(256)
hist_reduce_kernel(const double *__restrict__ partials, int n_blocks, int n_bins,
                   double *__restrict__ hist, double *__restrict__ hist_w2) {
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; // warp-uniform
    const int lane = threadIdx.x & 31;
    if (b >= 2 * n_bins) return;
    double s = 0.0;
    for (int k = lane; k < n_blocks; k += 32) s += __ldg(partials + (size_t)k * 2 * n_bins + b);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        if (b < n_bins) hist[b] = s;
        else if (hist_w2) hist_w2[b - n_bins] = s;
    }
}

// Batched variant for the fused template kernel: partials[container][block][2 n_bins] ->
// out[container][2][n_bins], same fixed summation order per value.
__global__ void __launch_bounds__(256)
hist_reduce_batch_kernel(const double *__restrict__ partials, int n_blocks, int n_bins, int n_containers,
                         double *__restrict__ out) {
    const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; // warp-uniform: (container, value)
    const int lane = threadIdx.x & 31;
    if (v >= n_containers * 2 * n_bins) return;
    const int c = v / (2 * n_bins), b = v - c * 2 * n_bins;
    const double *src = partials + (size_t)c * n_blocks * 2 * n_bins + b;
    double s = 0.0;
    for (int k = lane; k < n_blocks; k += 32) s += __ldg(src + (size_t)k * 2 * n_bins);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[v] = s;
}

// Fit-loop epilogue in ONE launch: the per-container reduction of hist_reduce_batch_kernel, the per-bin
// detector-systematics scale of discr_sys.hypersurfaces (hypersurfaces.py:219-243: errors *= s, weights =
// clip(weights * s, 0); here sum w *= max(s, 0)-clipped product and sum w^2 *= s^2), the MapSet sum over containers
// and mod_chi2 against the observed map (stats.py:651-695).  Blocks reduce their values as before; the LAST block to
// finish (device-scope arrival counter) owns the epilogue, so no second launch is needed and nothing synchronises.
__global__ void __launch_bounds__(256)
hist_reduce_chi2_kernel(const double *__restrict__ partials, int n_blocks, int n_bins, int n_containers,
                        const double *__restrict__ bin_scales, const double *__restrict__ observed,
                        double *__restrict__ out, double *__restrict__ total, double *__restrict__ chi2,
                        unsigned *__restrict__ arrive) {
    const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; // warp-uniform: (container, plane, bin)
    if (v < n_containers * 2 * n_bins) {
        const int c = v / (2 * n_bins), b = v - c * 2 * n_bins;
        reduce_container_value(partials + (size_t)c * n_blocks * 2 * n_bins, n_blocks, n_bins, b,
                               bin_scales ? bin_scales + (size_t)c * n_bins : nullptr, out + (size_t)c * 2 * n_bins);
    }
    __shared__ bool s_last;
    __shared__ double s_scratch[2 * PISAB_DET_MAX_BINS + 256];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(arrive, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) *arrive = 0; // the next launch on this stream starts from zero
    __threadfence();
    template_total_chi2(out, n_containers, n_bins, observed, total, chi2, s_scratch);
}

int hist_reduce_chi2(const double *d_partials, int n_blocks, int n_bins, int n_containers, const double *d_bin_scales,
                     const double *d_observed, double *d_out, double *d_total, double *d_chi2, unsigned *d_arrive,
                     cudaStream_t s) {
    const int warps = n_containers * 2 * n_bins;
    hist_reduce_chi2_kernel<<<(warps * 32 + 255) / 256, 256, 0, s>>>(d_partials, n_blocks, n_bins, n_containers,
                                                                      d_bin_scales, d_observed, d_out, d_total, d_chi2, d_arrive);
    note_launch();
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}

int hist_reduce_batch(const double *d_partials, int n_blocks, int n_bins, int n_containers, double *d_out,
                      cudaStream_t s) {
    const int warps = n_containers * 2 * n_bins;
    hist_reduce_batch_kernel<<<(warps * 32 + 255) / 256, 256, 0, s>>>(d_partials, n_blocks, n_bins, n_containers, d_out);
    note_launch();
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}

int hist_grid(int64_t n) {
    const int sms = sm_count() > 0 ? sm_count() : 148;
    int64_t want = (n + kHistBlock - 1) / kHistBlock;
    if (want < 1) want = 1;
    const int64_t cap = (int64_t)sms * 8;
    return (int)(want < cap ? want : cap);
}

int hist_reduce_partials(const double *d_partials, int n_blocks, int n_bins, double *d_hist,
                         double *d_hist_w2, cudaStream_t s) {
    hist_reduce_kernel<<<(2 * n_bins * 32 + 255) / 256, 256, 0, s>>>(d_partials, n_blocks, n_bins, d_hist, d_hist_w2);
    note_launch();
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}

// ---------------------------------------------------------------------------------------------
// lookup, fill_probs, apply weights, mod_chi2
// ---------------------------------------------------------------------------------------------
template <typename IO>
__global__ void __launch_bounds__(256)
lookup_kernel(const int32_t *__restrict__ index, const IO *__restrict__ flat_hist, int64_t n,
              int width, IO *__restrict__ out) {
    const int64_t total = n * width;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += stride) {
        const int64_t i = k / width;
        const int w = (int)(k - i * width);
        const int b = __ldg(index + i);
        out[k] = b >= 0 ? __ldg(flat_hist + (int64_t)b * width + w) : (IO)0;
    }
}

// width == 1 (every per-event scalar: prob_e, prob_mu, weights ...): no index arithmetic, 4 independent
// gathers per thread in flight.  HBM-bound: 4 B in + sizeof(IO) out per event.
template <typename IO>
__global__ void __launch_bounds__(256)
lookup1_kernel(const int32_t *__restrict__ index, const IO *__restrict__ flat_hist, int64_t n,
               IO *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        int b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) b[u] = __ldg(index + i + u * stride);
        IO v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = b[u] >= 0 ? __ldg(flat_hist + b[u]) : (IO)0;
#pragma unroll
        for (int u = 0; u < 4; ++u) out[i + u * stride] = v[u];
    }
    for (; i < n; i += stride) {
        const int b = __ldg(index + i);
        out[i] = b >= 0 ? __ldg(flat_hist + b) : (IO)0;
    }
}

template <typename IO>
__global__ void __launch_bounds__(256)
fill_probs_kernel(const IO *__restrict__ probability, int offset, int64_t n, IO *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = __ldg(probability + i * 9 + offset);
}

template <typename IO>
__global__ void __launch_bounds__(256)
apply_osc_weights_kernel(const IO *__restrict__ nu_flux, const IO *__restrict__ prob_e,
                         const IO *__restrict__ prob_mu, int64_t n, IO *__restrict__ weights) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        // prob3.py:622, evaluated in FTYPE without contraction like numpy
        const IO a = nu_flux[2 * i] * prob_e[i];
        const IO b = nu_flux[2 * i + 1] * prob_mu[i];
        if (sizeof(IO) == 8) weights[i] = (IO)__dmul_rn((double)weights[i], __dadd_rn((double)a, (double)b));
        else weights[i] = (IO)__fmul_rn((float)weights[i], __fadd_rn((float)a, (float)b));
    }
}

// stats.py:651-695 mod_chi2 = sum (N_obs - N_exp)^2 / (sigma^2 + N_exp), N_exp clipped at 1e-10
__global__ void __launch_bounds__(256)
mod_chi2_kernel(const double *__restrict__ expected, const double *__restrict__ expected_w2,
                const double *__restrict__ observed, int n_bins, double *__restrict__ out) {
    __shared__ double s[256];
    double acc = 0.0;
    for (int b = threadIdx.x; b < n_bins; b += blockDim.x) {
        const double e = fmax(expected[b], 1e-10);
        const double sig2 = expected_w2 ? expected_w2[b] : 0.0;
        const double d = observed[b] - e;
        acc += d * d / (sig2 + e);
    }
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) s[threadIdx.x] += s[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = s[0];
}

// One template -> one number: total map = sum over containers (container order), sigma^2 = sum of the
// containers' sum-w^2 maps (MapSet sum + sumw2 errors, hist.py:205-218), then mod_chi2 against `observed`.
// Optionally writes the summed map / sigma^2 (total[2][n_bins]).
__global__ void __launch_bounds__(256)
template_chi2_kernel(const double *__restrict__ hist, int n_containers, int n_bins,
                     const double *__restrict__ observed, double *__restrict__ total,
                     double *__restrict__ out) {
    __shared__ double s[256];
    double acc = 0.0;
    for (int b = threadIdx.x; b < n_bins; b += blockDim.x) {
        double e = 0.0, sig2 = 0.0;
        for (int c = 0; c < n_containers; ++c) {
            e += hist[((size_t)c * 2) * n_bins + b];
            sig2 += hist[((size_t)c * 2 + 1) * n_bins + b];
        }
        if (total) { total[b] = e; total[n_bins + b] = sig2; }
        e = fmax(e, 1e-10);
        const double d = observed[b] - e;
        acc += d * d / (sig2 + e);
    }
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) s[threadIdx.x] += s[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = s[0];
}

// one block per template of a scan: out[t] = mod_chi2(sum_c hist[t][c][0], sigma^2 = sum_c hist[t][c][1] | observed)
__global__ void __launch_bounds__(128)
template_chi2_batch_kernel(const double *__restrict__ hist, int n_containers, int n_bins,
                           const double *__restrict__ observed, double *__restrict__ out) {
    __shared__ double s[128];
    const double *mine = hist + (size_t)blockIdx.x * n_containers * 2 * n_bins;
    double acc = 0.0;
    for (int b = threadIdx.x; b < n_bins; b += blockDim.x) {
        double e = 0.0, sig2 = 0.0;
        for (int c = 0; c < n_containers; ++c) {
            e += mine[((size_t)c * 2) * n_bins + b];
            sig2 += mine[((size_t)c * 2 + 1) * n_bins + b];
        }
        e = fmax(e, 1e-10);
        const double d = observed[b] - e;
        acc += d * d / (sig2 + e);
    }
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int off = 64; off > 0; off >>= 1) {
        if (threadIdx.x < off) s[threadIdx.x] += s[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = s[0];
}

// aeff.aeff (pisa/stages/aeff/aeff.py:68-88): weights *= weighted_aeff * scale, evaluated in FTYPE like numpy does
// (the temporary `weighted_aeff * scale` is rounded to FTYPE before the in-place multiply).  24 B/event, HBM-bound.
template <typename IO>
__global__ void __launch_bounds__(256)
scale_weights_kernel(const IO *__restrict__ factor, double scale, int64_t n, IO *__restrict__ weights) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const IO s = (IO)scale;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (sizeof(IO) == 8) {
            const double t = factor ? __dmul_rn((double)__ldg(factor + i), (double)s) : (double)s;
            weights[i] = (IO)__dmul_rn((double)weights[i], t);
        } else {
            const float t = factor ? __fmul_rn((float)__ldg(factor + i), (float)s) : (float)s;
            weights[i] = (IO)__fmul_rn((float)weights[i], t);
        }
    }
}

// flat index on the joint binning (calc_mode + apply_mode of utils.hist, hist.py:69-84) from the two cached
// sub-indices: lifts the PISAB_MAX_DIMS limit of hist_index for the 5-dimensional true x reco layouts
__global__ void __launch_bounds__(256)
joint_index_kernel(const int32_t *__restrict__ a, const int32_t *__restrict__ b, int32_t size_b, int64_t n,
                   int32_t *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int32_t ia = __ldg(a + i), ib = __ldg(b + i);
        out[i] = (ia < 0 || ib < 0) ? -1 : ia * size_b + ib;
    }
}

// utils.hist with a binned calc_mode (hist.py:131-160): hist = (unc w) @ T, sumw2 = (unc w)^2 @ T,
// bin_unc2 = (unc^2 w) @ T with T = hist_transform [n_calc, n_out].  One block serves 32 output bins: thread
// (ox, cy) walks the calc bins cy, cy + 8, ... (coalesced rows of T), the 8 partial sums per output bin are added in
// fixed order -- bit-reproducible.  Setup-size arithmetic (n_calc x n_out ~ 5e6 products).
template <typename IO>
__global__ void __launch_bounds__(256)
hist_transform_kernel(const IO *__restrict__ w, const IO *__restrict__ unc, const IO *__restrict__ T, int n_calc,
                      int n_out, double *__restrict__ h, double *__restrict__ sumw2, double *__restrict__ bin_unc2) {
    __shared__ double s[3][8][33];
    const int ox = threadIdx.x & 31, cy = threadIdx.x >> 5;
    const int o = blockIdx.x * 32 + ox;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    if (o < n_out) {
        for (int c = cy; c < n_calc; c += 8) {
            const double wc = (double)__ldg(w + c), u = unc ? (double)__ldg(unc + c) : 1.0;
            const double t = (double)__ldg(T + (size_t)c * n_out + o);
            const double uw = u * wc;
            a0 = fma(uw, t, a0);
            a1 = fma(uw * uw, t, a1);
            a2 = fma(u * uw, t, a2);
        }
    }
    s[0][cy][ox] = a0; s[1][cy][ox] = a1; s[2][cy][ox] = a2;
    __syncthreads();
    if (cy < 3 && o < n_out) {
        double acc = 0.0;
        for (int k = 0; k < 8; ++k) acc += s[cy][k][ox];
        double *dst = cy == 0 ? h : (cy == 1 ? sumw2 : bin_unc2);
        if (dst) dst[o] = acc;
    }
}

static int ew_grid(int64_t n) {
    const int sms = sm_count() > 0 ? sm_count() : 148;
    int64_t want = (n + 255) / 256;
    if (want < 1) want = 1;
    const int64_t cap = (int64_t)sms * 8;
    return (int)(want < cap ? want : cap);
}

template <typename IO>
static int hist_index_impl(const pisab_binning_t *binning, const IO *const *d_coords, int64_t n,
                           int32_t *d_index, void *stream) {
    if (!binning || !d_coords || n < 0 || (n > 0 && !d_index)) { set_error("bad arguments"); return PISAB_ERR_ARG; }
    if (binning->n_dims < 1 || binning->n_dims > PISAB_MAX_DIMS) { set_error("n_dims outside [1,%d]", PISAB_MAX_DIMS); return PISAB_ERR_ARG; }
    BinningTable B = {};
    CoordPtrs<IO> C = {};
    B.n_dims = binning->n_dims;
    int64_t total = 1;
    for (int d = 0; d < B.n_dims; ++d) {
        B.kind[d] = binning->kind[d];
        B.n_bins[d] = binning->n_bins[d];
        if (B.n_bins[d] < 1) { set_error("dimension %d has no bins", d); return PISAB_ERR_ARG; }
        total *= B.n_bins[d];
        B.lo[d] = binning->lo[d];
        B.hi[d] = binning->hi[d];
        // norm = n / (hi - lo)  (translation.py:419,429-430)
        B.norm[d] = (double)B.n_bins[d] / (B.hi[d] - B.lo[d]);
        B.edges[d] = binning->d_edges[d];
        if (B.kind[d] == PISAB_DIM_EDGES && !B.edges[d]) { set_error("dimension %d: edges missing", d); return PISAB_ERR_ARG; }
        if (B.kind[d] != PISAB_DIM_EDGES && !(B.hi[d] > B.lo[d])) { set_error("dimension %d: empty range", d); return PISAB_ERR_ARG; }
        if (B.kind[d] == PISAB_DIM_LOG && !(B.lo[d] > 0.0)) { set_error("dimension %d: a logarithmic axis needs a positive domain", d); return PISAB_ERR_ARG; }
        C.p[d] = d_coords[d];
        if (!C.p[d]) { set_error("dimension %d: null sample", d); return PISAB_ERR_ARG; }
    }
    if (total > 2147483647LL) { set_error("too many bins"); return PISAB_ERR_ARG; }
    if (n == 0) return PISAB_OK;
    switch (B.n_dims) {
    case 1: hist_index_kernel<IO, 1><<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(B, C, n, d_index); break;
    case 2: hist_index_kernel<IO, 2><<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(B, C, n, d_index); break;
    case 3: hist_index_kernel<IO, 3><<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(B, C, n, d_index); break;
    default: hist_index_kernel<IO, 4><<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(B, C, n, d_index); break;
    }
    note_launch();
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}

template <typename IO>
static int hist_accumulate_impl(const int32_t *d_index, const IO *d_weights, int64_t n, int32_t n_bins,
                                double *d_hist, double *d_hist_w2, void *d_workspace,
                                int64_t workspace_bytes, void *stream) {
    if (n < 0 || n_bins < 1 || !d_hist || (n > 0 && !d_index)) { set_error("bad arguments"); return PISAB_ERR_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    if (!d_workspace || workspace_bytes < pisab_hist_workspace_bytes(n, n_bins)) {
        set_error("workspace too small: need %lld bytes", (long long)pisab_hist_workspace_bytes(n, n_bins));
        return PISAB_ERR_WORKSPACE;
    }
    if (n_bins > PISAB_DET_MAX_BINS) {
        // [bound (8 bytes, padded to 16)] [FixedAcc 2 x n_bins]
        unsigned long long *d_bound = (unsigned long long *)d_workspace;
        FixedAcc *d_acc = (FixedAcc *)((char *)d_workspace + 16);
        PISAB_CUDA_CHECK(cudaMemsetAsync(d_workspace, 0, 16 + sizeof(FixedAcc) * 2 * (size_t)n_bins, s));
        if (n > 0) {
            LaunchTimer t(s);
            if (d_weights) {
                absmax_kernel<IO><<<ew_grid(n), 256, 0, s>>>(d_weights, n, d_bound);
                note_launch();
            }
            hist_accumulate_fixed_kernel<IO><<<ew_grid(n), 256, 0, s>>>(d_index, d_weights, n, n_bins, d_bound, d_acc,
                                                                        d_hist_w2 != nullptr);
            note_launch();
        }
        fixed_finish_kernel<<<(n_bins + 255) / 256, 256, 0, s>>>(d_acc, n_bins, d_bound, d_weights != nullptr,
                                                                (double)(n > 0 ? n : 1), d_hist, d_hist_w2);
        note_launch();
        PISAB_CUDA_CHECK(cudaGetLastError());
        return PISAB_OK;
    }
    int grid;
    if (n_bins <= kSlotsMaxBins) {
        // replicated-slot kernel: R chosen so that a warp's bins stay <= 16 KB
#ifndef PISAB_HIST_U
#define PISAB_HIST_U 12
#endif
#ifndef PISAB_HIST_R
#define PISAB_HIST_R 8
#endif
        constexpr int U = PISAB_HIST_U, R0 = PISAB_HIST_R;
        const int R = n_bins <= 128 ? R0 : (n_bins <= 256 ? 4 : 2);
        const size_t smem = (size_t)(kHistBlock / 32) * n_bins * R * sizeof(double2);
        auto kernel = R == R0 ? hist_accumulate_slots_kernel<IO, R0, U>
                              : (R == 4 ? hist_accumulate_slots_kernel<IO, 4, U> : hist_accumulate_slots_kernel<IO, 2, U>);
        PISAB_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kHistBlock, smem) != cudaSuccess || occ < 1) occ = 1;
        if (occ > 8) occ = 8; // workspace bound (pisab_hist_workspace_bytes)
        const int sms = sm_count() > 0 ? sm_count() : 148;
        int64_t want = (n + kHistBlock - 1) / kHistBlock;
        if (want < 1) want = 1;
        grid = (int)(want < (int64_t)sms * occ ? want : (int64_t)sms * occ);
        LaunchTimer t(s);
        kernel<<<grid, kHistBlock, smem, s>>>(d_index, d_weights, n, n_bins, (double *)d_workspace);
        note_launch();
    } else {
        grid = hist_grid(n);
        const size_t smem = WarpHist::smem_bytes(kHistBlock, n_bins);
        if (smem > 48 * 1024)
            PISAB_CUDA_CHECK(cudaFuncSetAttribute(hist_accumulate_kernel<IO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LaunchTimer t(s);
        hist_accumulate_kernel<IO><<<grid, kHistBlock, smem, s>>>(d_index, d_weights, n, n_bins, (double *)d_workspace);
        note_launch();
    }
    PISAB_CUDA_CHECK(cudaGetLastError());
    return hist_reduce_partials((const double *)d_workspace, grid, n_bins, d_hist, d_hist_w2, s);
}

} // namespace pisab

using namespace pisab;

extern "C" {

int64_t pisab_hist_workspace_bytes(int64_t n, int32_t n_bins) {
    (void)n;
    const int sms = sm_count() > 0 ? sm_count() : 148;
    return (int64_t)sms * 16 * 2 * (int64_t)n_bins * (int64_t)sizeof(double); // up to 16 blocks per SM-slot
}

int64_t pisab_reweight_batch_workspace_bytes(int32_t n_containers, int32_t n_bins) {
    if (n_containers < 1) n_containers = 1;
    return (int64_t)n_containers * pisab_hist_workspace_bytes(0, n_bins);
}

int pisab_hist_index_f64(const pisab_binning_t *binning, const double *const *d_coords, int64_t n,
                         int32_t *d_index, void *stream) {
    return hist_index_impl<double>(binning, d_coords, n, d_index, stream);
}
int pisab_hist_index_f32(const pisab_binning_t *binning, const float *const *d_coords, int64_t n,
                         int32_t *d_index, void *stream) {
    return hist_index_impl<float>(binning, d_coords, n, d_index, stream);
}
int pisab_hist_accumulate_f64(const int32_t *d_index, const double *d_weights, int64_t n, int32_t n_bins,
                              double *d_hist, double *d_hist_w2, void *d_workspace,
                              int64_t workspace_bytes, void *stream) {
    return hist_accumulate_impl<double>(d_index, d_weights, n, n_bins, d_hist, d_hist_w2, d_workspace,
                                        workspace_bytes, stream);
}
int pisab_hist_accumulate_f32(const int32_t *d_index, const float *d_weights, int64_t n, int32_t n_bins,
                              double *d_hist, double *d_hist_w2, void *d_workspace,
                              int64_t workspace_bytes, void *stream) {
    return hist_accumulate_impl<float>(d_index, d_weights, n, n_bins, d_hist, d_hist_w2, d_workspace,
                                       workspace_bytes, stream);
}

#define PISAB_EW_CHECK(cond)                         \
    if (!(cond)) {                                   \
        set_error("bad arguments: " #cond);          \
        return PISAB_ERR_ARG;                        \
    }

} // extern "C" (reopened below)

extern "C" int64_t pisab_hist_plan_bytes(int64_t n, int32_t n_bins) {
    if (n < 0 || n_bins < 1 || n_bins > kPlanMaxBins) return 0; // not plannable: use pisab_hist_accumulate_*
    return (int64_t)PlanLayout(n > 0 ? n : 1, n_bins).bytes();
}

extern "C" int pisab_hist_plan_build(const int32_t *d_index, int64_t n, int32_t n_bins, void *d_plan, int64_t plan_bytes,
                          void *stream) {
    PISAB_EW_CHECK(n >= 0 && n_bins >= 1 && n_bins <= kPlanMaxBins && (n == 0 || d_index) && d_plan);
    const PlanLayout L(n > 0 ? n : 1, n_bins);
    if (plan_bytes < (int64_t)L.bytes()) { set_error("plan buffer too small: need %lld bytes", (long long)L.bytes()); return PISAB_ERR_WORKSPACE; }
    if (n == 0) return PISAB_OK;
    uint16_t *off = (uint16_t *)d_plan, *perm = (uint16_t *)((char *)d_plan + L.off_bytes);
    hist_plan_kernel<<<(unsigned)L.n_tiles, n_bins <= 128 ? 128 : 256, 0, (cudaStream_t)stream>>>(d_index, n, n_bins, off, L.off_stride, perm);
    note_launch();
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}

template <typename IO>
static int hist_planned_impl(const void *d_plan, const IO *d_weights, int64_t n, int32_t n_bins, double *d_hist,
                             double *d_hist_w2, void *d_workspace, int64_t workspace_bytes, void *stream) {
    if (n < 0 || n_bins < 1 || n_bins > kPlanMaxBins || !d_plan || !d_hist || (n > 0 && !d_weights)) { set_error("bad arguments"); return PISAB_ERR_ARG; }
    if (((uintptr_t)d_weights & 15) != 0) { set_error("planned histogram: weights must be 16-byte aligned"); return PISAB_ERR_ARG; }
    if (!d_workspace || workspace_bytes < pisab_hist_workspace_bytes(n, n_bins)) {
        set_error("workspace too small: need %lld bytes", (long long)pisab_hist_workspace_bytes(n, n_bins));
        return PISAB_ERR_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const PlanLayout L(n > 0 ? n : 1, n_bins);
    const uint16_t *off = (const uint16_t *)d_plan, *perm = (const uint16_t *)((const char *)d_plan + L.off_bytes);
    const size_t stage = (size_t)kPlanTile * sizeof(IO) + (size_t)kPlanTile * 2 + (size_t)L.off_stride * 2;
    const size_t smem = 2 * stage;
    auto kernel = n_bins <= 128   ? hist_planned_kernel<IO, 128, 1>
                  : n_bins <= 256 ? hist_planned_kernel<IO, 256, 1>
                  : n_bins <= 512 ? hist_planned_kernel<IO, 256, 2>
                                  : hist_planned_kernel<IO, 256, 4>;
    const int threads = n_bins <= 128 ? 128 : 256;
    PISAB_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem) != cudaSuccess || occ < 1) occ = 1;
    if (occ > 8) occ = 8; // workspace bound (pisab_hist_workspace_bytes)
    const int sms = sm_count() > 0 ? sm_count() : 148;
    int64_t grid = (int64_t)sms * occ;
    if (grid > L.n_tiles) grid = L.n_tiles;
    if (n == 0) grid = 0;
    if (grid > 0) {
        LaunchTimer t(s);
        kernel<<<(unsigned)grid, threads, smem, s>>>(off, L.off_stride, perm, d_weights, n, n_bins, (double *)d_workspace);
        note_launch();
        PISAB_CUDA_CHECK(cudaGetLastError());
    }
    return hist_reduce_partials((const double *)d_workspace, (int)grid, n_bins, d_hist, d_hist_w2, s);
}

template <typename IO>
static int hist_sorted_impl(const int32_t *d_perm, const int32_t *d_sorted_index, const IO *d_weights, int64_t n,
                            int32_t n_bins, double *d_hist, double *d_hist_w2, void *d_workspace, int64_t workspace_bytes,
                            void *stream) {
    if (n < 0 || n_bins < 1 || !d_hist || (n > 0 && (!d_perm || !d_sorted_index || !d_weights))) { set_error("bad arguments"); return PISAB_ERR_ARG; }
    const int64_t need = 16 + (int64_t)sizeof(FixedAcc) * 2 * n_bins;
    if (!d_workspace || workspace_bytes < need) { set_error("workspace too small: need %lld bytes", (long long)need); return PISAB_ERR_WORKSPACE; }
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long *d_bound = (unsigned long long *)d_workspace;
    FixedAcc *d_acc = (FixedAcc *)((char *)d_workspace + 16);
    PISAB_CUDA_CHECK(cudaMemsetAsync(d_workspace, 0, (size_t)need, s));
    if (n > 0) {
        LaunchTimer t(s);
        absmax_kernel<IO><<<ew_grid(n), 256, 0, s>>>(d_weights, n, d_bound);
        note_launch();
        hist_accumulate_sorted_kernel<IO><<<ew_grid(n), 256, 0, s>>>(d_perm, d_sorted_index, d_weights, n, n_bins, d_bound, d_acc);
        note_launch();
    }
    fixed_finish_kernel<<<(n_bins + 255) / 256, 256, 0, s>>>(d_acc, n_bins, d_bound, true, (double)(n > 0 ? n : 1), d_hist, d_hist_w2);
    note_launch();
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}

extern "C" {
int pisab_hist_accumulate_sorted_f64(const int32_t *d_perm, const int32_t *d_sorted_index, const double *d_weights,
                                     int64_t n, int32_t n_bins, double *d_hist, double *d_hist_w2, void *d_workspace,
                                     int64_t workspace_bytes, void *stream) {
    return hist_sorted_impl<double>(d_perm, d_sorted_index, d_weights, n, n_bins, d_hist, d_hist_w2, d_workspace, workspace_bytes, stream);
}
int pisab_hist_accumulate_sorted_f32(const int32_t *d_perm, const int32_t *d_sorted_index, const float *d_weights,
                                     int64_t n, int32_t n_bins, double *d_hist, double *d_hist_w2, void *d_workspace,
                                     int64_t workspace_bytes, void *stream) {
    return hist_sorted_impl<float>(d_perm, d_sorted_index, d_weights, n, n_bins, d_hist, d_hist_w2, d_workspace, workspace_bytes, stream);
}
int pisab_hist_accumulate_planned_f64(const void *d_plan, const double *d_weights, int64_t n, int32_t n_bins,
                                      double *d_hist, double *d_hist_w2, void *d_workspace, int64_t workspace_bytes,
                                      void *stream) {
    return hist_planned_impl<double>(d_plan, d_weights, n, n_bins, d_hist, d_hist_w2, d_workspace, workspace_bytes, stream);
}
int pisab_hist_accumulate_planned_f32(const void *d_plan, const float *d_weights, int64_t n, int32_t n_bins,
                                      double *d_hist, double *d_hist_w2, void *d_workspace, int64_t workspace_bytes,
                                      void *stream) {
    return hist_planned_impl<float>(d_plan, d_weights, n, n_bins, d_hist, d_hist_w2, d_workspace, workspace_bytes, stream);
}

int pisab_lookup_f64(const int32_t *d_index, const double *d_flat_hist, int64_t n, int32_t width,
                     double *d_out, void *stream) {
    PISAB_EW_CHECK(n >= 0 && width >= 1 && (n == 0 || (d_index && d_flat_hist && d_out)));
    if (n == 0) return PISAB_OK;
    if (width == 1) lookup1_kernel<double><<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(d_index, d_flat_hist, n, d_out);
    else lookup_kernel<double><<<ew_grid(n * width), 256, 0, (cudaStream_t)stream>>>(d_index, d_flat_hist, n, width, d_out);
    note_launch();
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}
int pisab_lookup_f32(const int32_t *d_index, const float *d_flat_hist, int64_t n, int32_t width,
                     float *d_out, void *stream) {
    PISAB_EW_CHECK(n >= 0 && width >= 1 && (n == 0 || (d_index && d_flat_hist && d_out)));
    if (n == 0) return PISAB_OK;
    if (width == 1) lookup1_kernel<float><<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(d_index, d_flat_hist, n, d_out);
    else lookup_kernel<float><<<ew_grid(n * width), 256, 0, (cudaStream_t)stream>>>(d_index, d_flat_hist, n, width, d_out);
    note_launch();
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}

int pisab_fill_probs_f64(const double *d_probability, int32_t initial_flav, int32_t flav, int64_t n,
                         double *d_out, void *stream) {
    PISAB_EW_CHECK(n >= 0 && initial_flav >= 0 && initial_flav < 3 && flav >= 0 && flav < 3 && (n == 0 || (d_probability && d_out)));
    if (n == 0) return PISAB_OK;
    fill_probs_kernel<double><<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(d_probability, initial_flav * 3 + flav, n, d_out);
    note_launch();
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}
int pisab_fill_probs_f32(const float *d_probability, int32_t initial_flav, int32_t flav, int64_t n,
                         float *d_out, void *stream) {
    PISAB_EW_CHECK(n >= 0 && initial_flav >= 0 && initial_flav < 3 && flav >= 0 && flav < 3 && (n == 0 || (d_probability && d_out)));
    if (n == 0) return PISAB_OK;
    fill_probs_kernel<float><<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(d_probability, initial_flav * 3 + flav, n, d_out);
    note_launch();
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}

int pisab_apply_osc_weights_f64(const double *d_nu_flux, const double *d_prob_e, const double *d_prob_mu,
                                int64_t n, double *d_weights, void *stream) {
    PISAB_EW_CHECK(n >= 0 && (n == 0 || (d_nu_flux && d_prob_e && d_prob_mu && d_weights)));
    if (n == 0) return PISAB_OK;
    apply_osc_weights_kernel<double><<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(d_nu_flux, d_prob_e, d_prob_mu, n, d_weights);
    note_launch();
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}
int pisab_apply_osc_weights_f32(const float *d_nu_flux, const float *d_prob_e, const float *d_prob_mu,
                                int64_t n, float *d_weights, void *stream) {
    PISAB_EW_CHECK(n >= 0 && (n == 0 || (d_nu_flux && d_prob_e && d_prob_mu && d_weights)));
    if (n == 0) return PISAB_OK;
    apply_osc_weights_kernel<float><<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(d_nu_flux, d_prob_e, d_prob_mu, n, d_weights);
    note_launch();
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}

int pisab_scale_weights_f64(const double *d_factor, double scale, int64_t n, double *d_weights, void *stream) {
    PISAB_EW_CHECK(n >= 0 && (n == 0 || d_weights));
    if (n == 0) return PISAB_OK;
    scale_weights_kernel<double><<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(d_factor, scale, n, d_weights);
    note_launch();
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}
int pisab_scale_weights_f32(const float *d_factor, double scale, int64_t n, float *d_weights, void *stream) {
    PISAB_EW_CHECK(n >= 0 && (n == 0 || d_weights));
    if (n == 0) return PISAB_OK;
    scale_weights_kernel<float><<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(d_factor, scale, n, d_weights);
    note_launch();
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}

int pisab_joint_index(const int32_t *d_index_a, const int32_t *d_index_b, int32_t size_b, int64_t n,
                      int32_t *d_out, void *stream) {
    PISAB_EW_CHECK(n >= 0 && size_b >= 1 && (n == 0 || (d_index_a && d_index_b && d_out)));
    if (n == 0) return PISAB_OK;
    joint_index_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(d_index_a, d_index_b, size_b, n, d_out);
    note_launch();
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}

int pisab_hist_transform_f64(const double *d_weights, const double *d_unc, const double *d_transform, int32_t n_calc,
                             int32_t n_out, double *d_hist, double *d_sumw2, double *d_bin_unc2, void *stream) {
    PISAB_EW_CHECK(n_calc >= 1 && n_out >= 1 && d_weights && d_transform && d_hist);
    hist_transform_kernel<double><<<(n_out + 31) / 32, 256, 0, (cudaStream_t)stream>>>(
        d_weights, d_unc, d_transform, n_calc, n_out, d_hist, d_sumw2, d_bin_unc2);
    note_launch();
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}
int pisab_hist_transform_f32(const float *d_weights, const float *d_unc, const float *d_transform, int32_t n_calc,
                             int32_t n_out, double *d_hist, double *d_sumw2, double *d_bin_unc2, void *stream) {
    PISAB_EW_CHECK(n_calc >= 1 && n_out >= 1 && d_weights && d_transform && d_hist);
    hist_transform_kernel<float><<<(n_out + 31) / 32, 256, 0, (cudaStream_t)stream>>>(
        d_weights, d_unc, d_transform, n_calc, n_out, d_hist, d_sumw2, d_bin_unc2);
    note_launch();
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}

int pisab_mod_chi2(const double *d_expected, const double *d_expected_w2, const double *d_observed,
                   int32_t n_bins, double *d_out, void *stream) {
    PISAB_EW_CHECK(n_bins >= 1 && d_expected && d_observed && d_out);
    mod_chi2_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(d_expected, d_expected_w2, d_observed, n_bins, d_out);
    note_launch();
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}

int pisab_template_chi2(const double *d_hist, int32_t n_containers, int32_t n_bins,
                        const double *d_observed, double *d_total, double *d_out, void *stream) {
    PISAB_EW_CHECK(n_bins >= 1 && n_containers >= 1 && d_hist && d_observed && d_out);
    template_chi2_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(d_hist, n_containers, n_bins, d_observed, d_total, d_out);
    note_launch();
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}

int pisab_template_chi2_batch(const double *d_hist, int32_t n_templates, int32_t n_containers, int32_t n_bins,
                              const double *d_observed, double *d_out, void *stream) {
    PISAB_EW_CHECK(n_templates >= 1 && n_bins >= 1 && n_containers >= 1 && d_hist && d_observed && d_out);
    template_chi2_batch_kernel<<<n_templates, 128, 0, (cudaStream_t)stream>>>(d_hist, n_containers, n_bins, d_observed, d_out);
    note_launch();
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}

} // extern "C"
