// exchange.cu -- the ONE exchange step of the path as a single kernel over NVLink peer memory.
//
// What is exchanged: the per-GPU histogram buffer [containers][2][n_bins] (24 KB for the 8x8x2 analysis binning) --
// or [templates][containers][2][n_bins] for a scan -- once per template.  The reference has no multi-GPU path; the
// round-1 form of this step was all_gather + clone + (N-1) eager adds + copy = ~10 launches per template, which is
// harmless next to an 11 ms step but dominates a 45 us analysis-size template.
//
// One process per GPU.  Every rank owns one exchange buffer (cudaMalloc, exported with cudaIpcGetMemHandle and mapped
// by all peers):   slots[2 parities][world][capacity] doubles  +  flags[2 parities][world] epochs.
// pisab_exchange_allreduce launches ONE kernel per rank:
//   1. push    : every thread stores its values into slot[parity][my rank] of EVERY rank's buffer (peer stores over
//                NVLink / NVSwitch; its own buffer included);
//   2. publish : __threadfence_system, a device-scope arrival counter; the last block of the launch then writes the
//                epoch into flag[parity][my rank] of every rank (release, system scope);
//   3. collect : every block waits until all `world` flags in its OWN buffer carry the epoch (acquire, system scope),
//                then sums the `world` slots IN RANK ORDER and writes the result -- bit-identical on every rank and
//                from run to run, whatever the arrival order.
// Two parities: a rank can be at most one template ahead of its slowest peer (it needs that peer's flag of template
// e+1, which the peer publishes only after it has collected template e), so the slots of template e are never
// overwritten before everybody has read them.  The waits are bounded (~4 s of clock64) and raise an error flag instead
// of hanging the GPU.  The grid is capped at one block per SM, so every block of the launch is resident while it waits.
#include <string.h>

#include "common.cuh"

namespace pisab {

constexpr int kMaxWorld = 16;

struct ExchangeDev {
    int rank, world;
    int64_t capacity;          // doubles per slot
    double *slots[kMaxWorld];  // base of every rank's buffer as mapped in THIS process
    unsigned *flags[kMaxWorld];
    unsigned *arrive;          // local arrival counter
    int *error;                // local error flag (1 = a wait timed out)
};

struct ExchangeCtx {
    ExchangeDev dev;
    void *local_base;
    void *peer_base[kMaxWorld];
    size_t bytes;
    unsigned epoch;
    bool connected;
};

static size_t slots_bytes(int world, int64_t capacity) { return (size_t)2 * world * capacity * sizeof(double); }
static size_t flags_offset(int world, int64_t capacity) { return (slots_bytes(world, capacity) + 255) / 256 * 256; }
static size_t total_bytes(int world, int64_t capacity) { return flags_offset(world, capacity) + 512; }

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(256)
exchange_allreduce_kernel(const __grid_constant__ ExchangeDev d, const double *__restrict__ in, double *__restrict__ out,
                          int64_t count, unsigned epoch) {
    const int parity = epoch & 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t my_slot = ((size_t)parity * d.world + d.rank) * d.capacity;
    // 1. push
    for (int64_t i = first; i < count; i += stride) {
        const double v = in[i];
        for (int q = 0; q < d.world; ++q) d.slots[q][my_slot + i] = v;
    }
    // 2. publish
    __threadfence_system();
    __syncthreads();
    __shared__ bool s_last;
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(d.arrive, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        if (threadIdx.x == 0) *d.arrive = 0; // the next launch on this stream starts from zero
        __threadfence_system();
        if (threadIdx.x < d.world) st_release_sys(d.flags[threadIdx.x] + parity * kMaxWorld + d.rank, epoch);
    }
    // 3. collect
    if (threadIdx.x < d.world) {
        const unsigned *f = d.flags[d.rank] + parity * kMaxWorld + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(f) != epoch) {
            if (clock64() - t0 > 8000000000LL) { // ~4 s: a peer died or never launched; do not hang the GPU
                *d.error = 1;
                break;
            }
            __nanosleep(100);
        }
    }
    __syncthreads();
    const double *mine = d.slots[d.rank] + (size_t)parity * d.world * d.capacity;
    for (int64_t i = first; i < count; i += stride) {
        double s = __ldcg(mine + i);
        for (int q = 1; q < d.world; ++q) s += __ldcg(mine + (size_t)q * d.capacity + i); // rank order: reproducible
        out[i] = s;
    }
}

// rank-ordered sum of gathered[world][count] (the fallback after a library all_gather): one launch
__global__ void __launch_bounds__(256)
sum_slots_kernel(const double *__restrict__ gathered, int world, int64_t count, double *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        double s = gathered[i];
        for (int q = 1; q < world; ++q) s += gathered[(size_t)q * count + i];
        out[i] = s;
    }
}

} // namespace pisab

using namespace pisab;

extern "C" {

int pisab_exchange_create(int32_t rank, int32_t world, int64_t capacity_doubles, void **ctx_out, unsigned char *handle_out) {
    if (!ctx_out || !handle_out || world < 1 || world > kMaxWorld || rank < 0 || rank >= world || capacity_doubles < 1) {
        set_error("exchange: bad arguments (world <= %d)", kMaxWorld);
        return PISAB_ERR_ARG;
    }
    ExchangeCtx *c = new ExchangeCtx();
    memset(c, 0, sizeof(*c));
    c->bytes = total_bytes(world, capacity_doubles);
    PISAB_CUDA_CHECK(cudaMalloc(&c->local_base, c->bytes));
    PISAB_CUDA_CHECK(cudaMemset(c->local_base, 0, c->bytes));
    PISAB_CUDA_CHECK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    PISAB_CUDA_CHECK(cudaIpcGetMemHandle(&h, c->local_base));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle_out, &h, 64);
    c->dev.rank = rank;
    c->dev.world = world;
    c->dev.capacity = capacity_doubles;
    c->epoch = 0;
    *ctx_out = c;
    return PISAB_OK;
}

int pisab_exchange_connect(void *ctx, const unsigned char *all_handles) {
    ExchangeCtx *c = (ExchangeCtx *)ctx;
    if (!c || !all_handles) { set_error("exchange: bad arguments"); return PISAB_ERR_ARG; }
    const int world = c->dev.world;
    const size_t foff = flags_offset(world, c->dev.capacity);
    for (int q = 0; q < world; ++q) {
        void *base = c->local_base;
        if (q != c->dev.rank) {
            cudaIpcMemHandle_t h;
            memcpy(&h, all_handles + (size_t)q * 64, 64);
            PISAB_CUDA_CHECK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
        }
        c->peer_base[q] = base;
        c->dev.slots[q] = (double *)base;
        c->dev.flags[q] = (unsigned *)((char *)base + foff);
    }
    // arrival counter and error flag live behind the flags of the local buffer (2 * kMaxWorld epochs = 128 bytes)
    c->dev.arrive = (unsigned *)((char *)c->local_base + foff + 256);
    c->dev.error = (int *)((char *)c->local_base + foff + 320);
    c->connected = true;
    return PISAB_OK;
}

int pisab_exchange_allreduce(void *ctx, double *d_buf, int64_t count, void *stream) {
    ExchangeCtx *c = (ExchangeCtx *)ctx;
    if (!c || !c->connected || !d_buf || count < 1) { set_error("exchange: not connected or bad arguments"); return PISAB_ERR_ARG; }
    if (count > c->dev.capacity) { set_error("exchange: %lld values exceed the capacity of %lld", (long long)count, (long long)c->dev.capacity); return PISAB_ERR_WORKSPACE; }
    c->epoch += 1;
    if (c->epoch == 0) c->epoch = 1;
    const int sms = sm_count() > 0 ? sm_count() : 148;
    int64_t grid = (count + 255) / 256;
    if (grid > sms) grid = sms; // every block must be resident while it waits for the peers
    exchange_allreduce_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(c->dev, d_buf, d_buf, count, c->epoch);
    note_launch();
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}

/* 0 = ok, 1 = a wait for a peer timed out at some point (synchronises the device). */
int pisab_exchange_status(void *ctx) {
    ExchangeCtx *c = (ExchangeCtx *)ctx;
    if (!c || !c->connected) return -1;
    int err = 0;
    if (cudaDeviceSynchronize() != cudaSuccess) return -2;
    if (cudaMemcpy(&err, c->dev.error, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
    return err;
}

/* Unmap the peers' buffers (step 1 of an orderly shutdown: every rank disconnects, the host synchronises the ranks,
 * then every rank destroys -- an exported buffer must not be freed while a peer still maps it). */
int pisab_exchange_disconnect(void *ctx) {
    ExchangeCtx *c = (ExchangeCtx *)ctx;
    if (!c) return PISAB_OK;
    cudaDeviceSynchronize();
    for (int q = 0; q < c->dev.world; ++q)
        if (c->connected && q != c->dev.rank && c->peer_base[q]) { cudaIpcCloseMemHandle(c->peer_base[q]); c->peer_base[q] = nullptr; }
    c->connected = false;
    return PISAB_OK;
}

int pisab_exchange_destroy(void *ctx) {
    ExchangeCtx *c = (ExchangeCtx *)ctx;
    if (!c) return PISAB_OK;
    pisab_exchange_disconnect(ctx);
    if (c->local_base) cudaFree(c->local_base);
    delete c;
    return PISAB_OK;
}

int pisab_sum_slots(const double *d_gathered, int32_t world, int64_t count, double *d_out, void *stream) {
    if (!d_gathered || !d_out || world < 1 || count < 0) { set_error("sum_slots: bad arguments"); return PISAB_ERR_ARG; }
    if (count == 0) return PISAB_OK;
    const int sms = sm_count() > 0 ? sm_count() : 148;
    int64_t grid = (count + 255) / 256;
    if (grid > (int64_t)sms * 8) grid = (int64_t)sms * 8;
    sum_slots_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(d_gathered, world, count, d_out);
    note_launch();
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}

} // extern "C"
