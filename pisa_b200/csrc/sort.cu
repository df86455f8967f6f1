// sort.cu -- setup-time ordering of events by a small integer key (stable LSD radix sort, cub::DeviceRadixSort).
//
// Two users, both in setup_function-like code that runs once per event sample, never per template:
//   * the order that groups events by the number of Earth shells they cross (prob3.setup_function computes its layer
//     arrays once as well, pisa/stages/osc/prob3.py:406-409), optionally with the bin index as secondary key;
//   * the sorted plan of a histogram with more bins than fit in shared memory (hist.cu, pisab_hist_accumulate_sorted_*).
// Stability matters: the permutation fixes the summation order of the histograms, hence their bits.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace pisab {

__global__ void __launch_bounds__(256) iota_kernel(int32_t *__restrict__ out, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (int32_t)i;
}

static size_t cub_temp_bytes(int64_t n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int32_t *)nullptr, (int32_t *)nullptr, (const int32_t *)nullptr,
                                    (int32_t *)nullptr, (int)n);
    return (bytes + 255) / 256 * 256;
}
static size_t round256(size_t b) { return (b + 255) / 256 * 256; }

} // namespace pisab

using namespace pisab;

extern "C" {

int64_t pisab_sort_workspace_bytes(int64_t n) {
    if (n < 1) n = 1;
    if (n > 2147483647LL) return 0;
    return (int64_t)(cub_temp_bytes(n) + 2 * round256((size_t)n * 4));
}

int pisab_sort_order_i32(const int32_t *d_keys, int64_t n, int32_t key_bits, int32_t descending, int32_t *d_order,
                         int32_t *d_sorted_keys, void *d_workspace, int64_t workspace_bytes, void *stream) {
    if (n < 0 || n > 2147483647LL || (n > 0 && (!d_keys || !d_order)) || key_bits < 0 || key_bits > 31) {
        set_error("sort: bad arguments (non-negative int32 keys, at most 2^31-1 of them, key_bits in 0..31)");
        return PISAB_ERR_ARG;
    }
    if (n == 0) return PISAB_OK;
    if (!d_workspace || workspace_bytes < pisab_sort_workspace_bytes(n)) {
        set_error("sort: workspace too small: need %lld bytes", (long long)pisab_sort_workspace_bytes(n));
        return PISAB_ERR_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    size_t temp = cub_temp_bytes(n);
    char *base = (char *)d_workspace;
    int32_t *d_iota = (int32_t *)(base + temp);
    int32_t *d_keys_out = d_sorted_keys ? d_sorted_keys : (int32_t *)(base + temp + round256((size_t)n * 4));
    const int sms = sm_count() > 0 ? sm_count() : 148;
    const int64_t want = (n + 255) / 256;
    iota_kernel<<<(unsigned)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8), 256, 0, s>>>(d_iota, n);
    note_launch();
    const int end_bit = key_bits ? key_bits : 31; // keys are non-negative
    cudaError_t e = descending
        ? cub::DeviceRadixSort::SortPairsDescending(d_workspace, temp, d_keys, d_keys_out, d_iota, d_order, (int)n, 0, end_bit, s)
        : cub::DeviceRadixSort::SortPairs(d_workspace, temp, d_keys, d_keys_out, d_iota, d_order, (int)n, 0, end_bit, s);
    PISAB_CUDA_CHECK(e);
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}

} // extern "C"
