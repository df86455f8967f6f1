// layers.cu -- Layers.calcLayers / extCalcLayers (layers.py:38-169) as a stand-alone kernel.
// Parity / API path: the propagation kernels evaluate the same geometry in registers and never
// materialise these [n, max_layers] arrays.
#include "prob3_device.cuh"

namespace pisab {

// Rows are [max_layers] long, so a thread writing its own row touches one sector per store.  Each warp
// therefore builds its 32 rows in shared memory and then copies the 32 * max_layers contiguous output elements with
// coalesced streaming stores.  The tile holds the distances as IO ([32][row], row stride odd in elements:
// conflict-free) and, instead of the densities, the one-byte SHELL index of every slot ([32][row_b] bytes, row stride
// odd in 32-bit words): the density is looked up in the Earth table during the copy-out.  8.4 KB per warp in FP64
// instead of 14.8 KB doubles the resident warps (24 per SM); the copy-out walks (row, column) incrementally (the
// first version divided by max_layers per element and reached 2.16 TB/s = 33 % of HBM).
__host__ __device__ inline int layers_row_stride(int max_layers) { return max_layers | 1; }
__host__ __device__ inline int layers_row_bytes(int max_layers) { return 4 * (((max_layers + 3) / 4) | 1); }
constexpr int kLayersWarps = 8;
constexpr unsigned char kNoShell = 255;

template <typename IO, bool FULL_ROWS>
__global__ void __launch_bounds__(32 * kLayersWarps)
layers_kernel(const __grid_constant__ EarthTable earth, int max_layers, int rows_arg, const IO *__restrict__ coszen,
              int64_t n, IO *__restrict__ densities, IO *__restrict__ distances,
              int32_t *__restrict__ n_layers) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ EarthTable E;
    {
        const int *src = reinterpret_cast<const int *>(&earth);
        int *dst = reinterpret_cast<int *>(&E);
        for (int i = threadIdx.x; i < (int)(sizeof(EarthTable) / 4); i += blockDim.x) dst[i] = src[i];
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = layers_row_stride(max_layers), row_b = layers_row_bytes(max_layers);
    const int rows = FULL_ROWS ? 32 : rows_arg; // (a compile-time 32 for the usual models: no predicates, no idle lanes)
    // [warps][rows][row] distances (IO), then [warps][rows][row_b] shell indices (bytes); `rows` (<= 32) directions per
    // warp and pass: 32 for the usual Earth models, fewer for very deep ones (PREM_59layer: 122 slots) so that the tile
    // does not cost the occupancy -- the lanes beyond `rows` only help with the copy-out
    const int n_warps = blockDim.x >> 5;
    IO *tile_dis = reinterpret_cast<IO *>(s_raw) + (size_t)warp * rows * row;
    unsigned char *tile_sh = s_raw + (size_t)n_warps * rows * row * sizeof(IO) + (size_t)warp * rows * row_b;
    const int64_t stride = (int64_t)gridDim.x * n_warps * rows;
    // warp-uniform trip count: every lane takes part in the copy-out of its warp's tile
    for (int64_t first = ((int64_t)blockIdx.x * n_warps + warp) * rows; first < n; first += stride) {
        const int64_t i = first + lane;
        const bool live = lane < rows && i < n;
        const IO czs = live ? __ldg(coszen + i) : (IO)1;
        const double cz = (double)czs;
        // (idle lanes compute the trivial down-going case into row 0's shadow: they write nothing, see `put`)
        unsigned char *sh = tile_sh + (lane < rows ? lane : 0) * row_b;
        IO *dis = tile_dis + (lane < rows ? lane : 0) * row;
        const bool writer = FULL_ROWS || lane < rows;
        const int idx = E.idx_first_inner;
        const double base = __dmul_rn(-E.r_det, cz);
        int count = 0, slot = 0;
        auto put = [&](int shell, double seg) {
            if (writer) {
                sh[slot] = seg > 0.0 ? (unsigned char)shell : kNoShell;
                dis[slot] = (IO)seg;
            }
            count += seg > 0.0;
            ++slot;
        };
        if (!(cz < E.limit[idx])) {
            // layers.py:94-103 (`coszen**2.` is float64 in both FTYPE modes)
            const double cz2 = __dmul_rn(cz, cz);
            double l_cur = __dadd_rn(base, shell_root(E.rd2, cz2, E.rj2[0]));
            for (int j = 0; j < idx; ++j) {
                const double l_next = (j + 1 < idx) ? __dadd_rn(base, shell_root(E.rd2, cz2, E.rj2[j + 1])) : 0.0;
                put(j, __dsub_rn(l_cur, l_next));
                l_cur = l_next;
            }
            if (writer)
                for (; slot < E.n_radii && slot < max_layers; ++slot) { sh[slot] = kNoShell; dis[slot] = (IO)0; }
        } else {
            // layers.py:105-159; `coszen**2` (int exponent) stays in FTYPE under numba's typing
            const double cz2 = sizeof(IO) == 4 ? (double)__fmul_rn((float)czs, (float)czs) : __dmul_rn(cz, cz);
            int K = idx; // number of crossed shells
            while (K < E.n_radii && E.limit[K] > cz) ++K;
            // inbound: shells 0 .. K-1 into slots 0 .. K-1; outbound: shells K-2 .. 1, segments s_{j+1} - s_j
            // (s_j = base - sqrt_j the small root, s_{idx-1} := 0) into slots K .. 2K-3.  Both sides of shell j need
            // the same two square roots, so the outbound segment of shell j is written (to slot 2K-2-j) in the inbound
            // iteration that has them -- the same operations on the same values as the reference's second loop.
            double sq_j = shell_root(E.rd2, cz2, E.rj2[0]);
            double l_cur = __dadd_rn(base, sq_j);
            for (int j = 0; j < K; ++j) {
                double seg;
                if (j + 1 < K) {
                    const double sq_next = shell_root(E.rd2, cz2, E.rj2[j + 1]);
                    const double l_next = __dadd_rn(base, sq_next);
                    seg = __dsub_rn(l_cur, l_next);
                    l_cur = l_next;
                    if (j >= 1) {
                        const double s_hi = __dsub_rn(base, sq_next);
                        const double s_lo = j >= idx ? __dsub_rn(base, sq_j) : 0.0;
                        const double out = __dsub_rn(s_hi, s_lo);
                        const int o = 2 * K - 2 - j;
                        if (writer) {
                            sh[o] = out > 0.0 ? (unsigned char)j : kNoShell;
                            dis[o] = (IO)out;
                        }
                        count += out > 0.0;
                    }
                    sq_j = sq_next;
                } else {
                    seg = __dsub_rn(l_cur, __dsub_rn(base, sq_j));
                }
                put(j, seg);
            }
            slot = 2 * K - 2 > slot ? 2 * K - 2 : slot;
        }
        if (writer)
            for (; slot < max_layers; ++slot) { sh[slot] = kNoShell; dis[slot] = (IO)0; }
        if (n_layers && live) n_layers[i] = count;
        __syncwarp();
        const int64_t rows_left = n - first;
        const int total = (int)(rows_left < rows ? rows_left : rows) * max_layers;
        IO *out_den = densities + first * max_layers, *out_dis = distances + first * max_layers;
        // element k = r * max_layers + c of the warp's contiguous output block; (r, c) advanced by 32 per step
        int r = lane / max_layers, c = lane - r * max_layers;
        const int dr = 32 / max_layers, dc = 32 - dr * max_layers;
        for (int k = lane; k < total; k += 32) {
            const unsigned char shell = tile_sh[r * row_b + c];
            __stcs(out_den + k, shell == kNoShell ? (IO)0 : (IO)E.rho[shell]);
            __stcs(out_dis + k, tile_dis[r * row + c]);
            r += dr;
            c += dc;
            if (c >= max_layers) { c -= max_layers; ++r; }
        }
        __syncwarp();
    }
}

template <typename IO>
__global__ void __launch_bounds__(256)
layer_count_kernel(const __grid_constant__ EarthTable earth, const IO *__restrict__ coszen, int64_t n,
                   int32_t *__restrict__ count) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double cz = (double)__ldg(coszen + i);
        int k = 0;
        for (int j = 0; j < earth.n_radii; ++j) k += earth.limit[j] > cz;
        count[i] = k;
    }
}

template <typename IO>
static int layer_count_impl(const pisab_earth_t *earth, const IO *d_coszen, int64_t n, int32_t *d_count,
                            void *stream) {
    if (n < 0 || (n > 0 && (!d_coszen || !d_count))) { set_error("bad arguments"); return PISAB_ERR_ARG; }
    EarthTable et;
    int rc = build_earth_table(earth, &et);
    if (rc) return rc;
    if (n == 0) return PISAB_OK;
    const int sms = sm_count() > 0 ? sm_count() : 148;
    int64_t want = (n + 255) / 256;
    const int grid = (int)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);
    layer_count_kernel<IO><<<grid, 256, 0, (cudaStream_t)stream>>>(et, d_coszen, n, d_count);
    note_launch();
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}

template <typename IO>
static int layers_impl(const pisab_earth_t *earth, const IO *d_coszen, int64_t n, IO *d_densities,
                       IO *d_distances, int32_t *d_n_layers, void *stream) {
    if (n < 0 || (n > 0 && (!d_coszen || !d_densities || !d_distances))) { set_error("bad arguments"); return PISAB_ERR_ARG; }
    EarthTable et;
    int rc = build_earth_table(earth, &et);
    if (rc) return rc;
    if (earth->max_layers < 2 * earth->n_radii - 2) { set_error("max_layers too small"); return PISAB_ERR_ARG; }
    if (n == 0) return PISAB_OK;
    const int sms = sm_count() > 0 ? sm_count() : 148;
    // 32 rows per warp: distances as IO + one byte of shell index per slot; as many warps per block (<= kLayersWarps)
    // as ~200 KB of shared memory hold (PREM_59layer: 122 slots)
    // rows per warp and pass: 32, halved until a block of kLayersWarps warps stays below ~72 KB (three blocks per SM)
    const size_t per_row = layers_row_stride(earth->max_layers) * sizeof(IO) + layers_row_bytes(earth->max_layers);
    int rows = 32;
    while (rows > 4 && (size_t)rows * per_row * kLayersWarps > 72 * 1024) rows /= 2;
    const size_t per_warp = (size_t)rows * per_row;
    const int warps = kLayersWarps;
    if (et.n_radii >= (int)kNoShell) { set_error("too many shells"); return PISAB_ERR_UNSUPPORTED; }
    const size_t smem = per_warp * warps;
    int64_t want = (n + rows * warps - 1) / (rows * warps);
    int occ = 0;
    auto kernel = rows == 32 ? layers_kernel<IO, true> : layers_kernel<IO, false>;
    PISAB_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 32 * warps, smem) != cudaSuccess || occ < 1) occ = 1;
    const int grid = (int)(want < (int64_t)sms * occ * 4 ? want : (int64_t)sms * occ * 4);
    kernel<<<grid, 32 * warps, smem, (cudaStream_t)stream>>>(et, earth->max_layers, rows, d_coszen, n, d_densities,
                                                             d_distances, d_n_layers);
    note_launch();
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}

} // namespace pisab

extern "C" {
int pisab_layers_calc_f64(const pisab_earth_t *earth, const double *d_coszen, int64_t n,
                          double *d_densities, double *d_distances, int32_t *d_n_layers, void *stream) {
    return pisab::layers_impl<double>(earth, d_coszen, n, d_densities, d_distances, d_n_layers, stream);
}
int pisab_layer_count_f64(const pisab_earth_t *earth, const double *d_coszen, int64_t n, int32_t *d_count,
                          void *stream) {
    return pisab::layer_count_impl<double>(earth, d_coszen, n, d_count, stream);
}
int pisab_layer_count_f32(const pisab_earth_t *earth, const float *d_coszen, int64_t n, int32_t *d_count,
                          void *stream) {
    return pisab::layer_count_impl<float>(earth, d_coszen, n, d_count, stream);
}
int pisab_layers_calc_f32(const pisab_earth_t *earth, const float *d_coszen, int64_t n, float *d_densities,
                          float *d_distances, int32_t *d_n_layers, void *stream) {
    return pisab::layers_impl<float>(earth, d_coszen, n, d_densities, d_distances, d_n_layers, stream);
}
}
