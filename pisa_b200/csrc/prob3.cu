// prob3.cu -- propagation kernels + their C-ABI entry points (include/pisa_b200.h).
#include <math.h>
#include <string.h>

#include <mutex>

#include <vector>

#include "flux_device.cuh"
#include "hist_device.cuh"
#include "prob3_decay.cuh"
#include "prob3_walk.cuh"

namespace pisab {

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
#ifndef PISAB_BLOCK
#define PISAB_BLOCK 256
#endif
#ifndef PISAB_MIN_BLOCKS
#define PISAB_MIN_BLOCKS 2
#endif
constexpr int kBlock = PISAB_BLOCK;

template <typename IO>
__device__ __forceinline__ double ld(const IO *p, int64_t i) { return (double)__ldg(p + i); }

__device__ __forceinline__ void copy_tables(const OscTable &osc, const EarthTable &earth,
                                            OscTable *s_osc, EarthTable *s_earth) {
    const double *src = reinterpret_cast<const double *>(&osc);
    double *dst = reinterpret_cast<double *>(s_osc);
    for (int i = threadIdx.x; i < (int)(sizeof(OscTable) / 8); i += blockDim.x) dst[i] = src[i];
    if (s_earth) {
        const int *srci = reinterpret_cast<const int *>(&earth);
        int *dsti = reinterpret_cast<int *>(s_earth);
        for (int i = threadIdx.x; i < (int)(sizeof(EarthTable) / 4); i += blockDim.x) dsti[i] = srci[i];
    }
    __syncthreads();
}

__device__ __forceinline__ void copy_earth(const EarthTable &earth, EarthTable *s_earth) {
    const int *srci = reinterpret_cast<const int *>(&earth);
    int *dsti = reinterpret_cast<int *>(s_earth);
    for (int i = threadIdx.x; i < (int)(sizeof(EarthTable) / 4); i += blockDim.x) dsti[i] = srci[i];
    __syncthreads();
}

// ---- per-thread asynchronous staging (LDGSTS): global -> shared without touching registers ----
template <int BYTES>
__device__ __forceinline__ void cp_async(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(s), "l"(gmem), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// One thread per event, layers computed in-kernel from coszen (osc.prob3 compute_function).
//   FULL: probability[n,3,3] ; otherwise prob_e / prob_mu of the event's final flavour.
// Same layout as the fused kernel: propagation state and h0 in per-thread shared-memory columns, the
// next event's energy / coszen staged by cp.async while the current one is propagated.
// MP = the FP32 mode's mixed-precision arithmetic (prob3_mp.cuh): state and per-event Hamiltonian in registers, so
// the dynamic shared memory shrinks to the two staging slots per thread.
#ifndef PISAB_MP_MIN_BLOCKS
#define PISAB_MP_MIN_BLOCKS 3
#endif
template <bool FULL>
static size_t earth_smem_bytes(size_t io_bytes, bool std_matter, bool mp = false) {
    if (mp) return (size_t)(FULL ? PropagatorSmemF<3, 3>::kSlots : PropagatorSmemF<1, 2>::kSlots) * kBlock * sizeof(float2) +
                   2 * (size_t)kBlock * io_bytes;
    const size_t doubles = (size_t)((FULL ? PropagatorSmem<3, 3>::kDoubles : PropagatorSmem<1, 2>::kDoubles) +
                                    (std_matter ? H0Smem<true>::kDoubles : H0Smem<false>::kDoubles)) * kBlock;
    return doubles * sizeof(double) + 2 * (size_t)kBlock * io_bytes;
}

template <typename IO, bool FULL, bool STD, bool MP = false>
__global__ void __launch_bounds__(kBlock, MP ? PISAB_MP_MIN_BLOCKS : PISAB_MIN_BLOCKS)
prob3_earth_kernel(const __grid_constant__ OscTable osc, const __grid_constant__ EarthTable earth,
                   int nubar, const int32_t *__restrict__ d_nubar, int flav,
                   const int32_t *__restrict__ d_flav, const IO *__restrict__ energy,
                   const IO *__restrict__ coszen, const int32_t *__restrict__ order, int64_t n,
                   IO *__restrict__ probability, IO *__restrict__ prob_e, IO *__restrict__ prob_mu) {
    constexpr int NR = FULL ? 3 : 1, NC = FULL ? 3 : 2;
    extern __shared__ __align__(16) double s_dyn_earth[];
    __shared__ EarthTable s_earth;
    double2(*s_state)[kBlock] = reinterpret_cast<double2(*)[kBlock]>(s_dyn_earth);
    double(*s_h0)[kBlock] = reinterpret_cast<double(*)[kBlock]>(s_dyn_earth + PropagatorSmem<NR, NC>::kDoubles * kBlock);
    float2(*s_statef)[kBlock] = reinterpret_cast<float2(*)[kBlock]>(s_dyn_earth); // FP32 mode: float2 state columns
    IO *s_e = MP ? reinterpret_cast<IO *>(&s_statef[PropagatorSmemF<NR, NC>::kSlots][0])
                 : reinterpret_cast<IO *>(&s_h0[H0Smem<STD>::kDoubles][0]);
    IO *s_cz = s_e + kBlock;
    copy_earth(earth, &s_earth);
    const int tid = threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + tid;
    // `order` (optional) lists the events grouped by number of crossed shells so that the 32
    // lanes of a warp walk the same number of layers; results go back to the event's own slot
    auto event_of = [&](int64_t t) -> int { return t < n ? (order ? __ldg(order + t) : (int)t) : -1; };
    int i_cur = event_of(first), i_next = event_of(first + stride);
    if (i_cur >= 0) { s_e[tid] = __ldg(energy + i_cur); s_cz[tid] = __ldg(coszen + i_cur); }
    for (int64_t t = first; t < n; t += stride) {
        const int i_nn = event_of(t + 2 * stride);
        const int64_t i = i_cur;
        const double e = (double)s_e[tid], cz = (double)s_cz[tid];
        if (i_next >= 0) {
            cp_async<sizeof(IO)>(&s_e[tid], energy + i_next);
            cp_async<sizeof(IO)>(&s_cz[tid], coszen + i_next);
        }
        cp_async_commit();
        const int nb = d_nubar ? __ldg(d_nubar + i) : nubar;
        const int fl = d_flav ? __ldg(d_flav + i) : flav;
        const double inv_e = rcp_fast(e);
        const Herm3 hh = herm_axpy(nb > 0 ? inv_e : -inv_e, osc.hv[0], osc.lr); // hv[1] = -hv[0]
        auto emit = [&](const auto &P) {
            if (FULL) {
                IO *o = probability + i * 9;
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) o[a * 3 + b] = (IO)P.prob(b, a); // P(a->b) = |A[b][a]|^2
            } else {
                prob_e[i] = (IO)P.prob(0, 0);
                prob_mu[i] = (IO)P.prob(0, 1);
            }
        };
        if constexpr (MP) {
            H0MP<STD> h0;
            h0.init(hh);
            PropagatorSmemF<NR, NC> P{&s_statef[0][tid], kBlock};
            propagate_earth<NR, NC, STD>(h0, osc, s_earth, cz, inv_e, nb, FULL ? 0 : fl, P);
            emit(P);
        } else {
            H0Smem<STD> h0{&s_h0[0][tid], kBlock};
            h0.store(hh);
            if (STD) h0.set_poly(hh);
            PropagatorSmem<NR, NC> P{&s_state[0][tid], kBlock};
            propagate_earth<NR, NC, STD>(h0, osc, s_earth, cz, inv_e, nb, FULL ? 0 : fl, P);
            emit(P);
        }
        cp_async_wait_all();
        i_cur = i_next;
        i_next = i_nn;
    }
}

// Explicit layer arrays: drop-in for propagate_array (numba_osc_hostfuncs.py:60-70),
// including the reference's layer cache rule (numba_osc_kernels.py:230-249): layer i reuses
// the matrix of the LAST earlier layer j whose density and distance both differ by < 1e-5;
// since that matrix was itself either computed or copied, the chain is followed to the layer
// that was actually computed and its (rho, d) are used.
template <typename IO>
__global__ void __launch_bounds__(kBlock)
prob3_layers_kernel(const __grid_constant__ OscTable osc, int nubar,
                    const int32_t *__restrict__ d_nubar, const IO *__restrict__ energy,
                    const IO *__restrict__ densities, const IO *__restrict__ distances, int64_t n,
                    int n_layers, IO *__restrict__ probability) {
    __shared__ OscTable s_osc;
    copy_tables(osc, *reinterpret_cast<const EarthTable *>(&osc), &s_osc, nullptr);
    const double T_SCALE = kTab[18];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double e = ld(energy, i);
        const int nb = d_nubar ? __ldg(d_nubar + i) : nubar;
        const double inv_e = rcp_fast(e);
        const Herm3 h0 = herm_axpy(inv_e, s_osc.hv[nb > 0 ? 0 : 1], s_osc.lr);
        const IO *rho = densities + i * n_layers;
        const IO *dist = distances + i * n_layers;
        Cplx M[3][3];
        bool first = true;
        for (int l = 0; l < n_layers; ++l) {
            IO d = __ldg(dist + l);
            if (!(d > (IO)0)) continue;
            IO r = __ldg(rho + l);
            // resolve the cache chain
            int src = l;
            for (;;) {
                int hit = -1;
                const IO rs = __ldg(rho + src), ds = __ldg(dist + src);
                for (int j = 0; j < src; ++j)
                    if (fabs((double)(__ldg(rho + j) - rs)) < 1e-5 && fabs((double)(__ldg(dist + j) - ds)) < 1e-5 &&
                        __ldg(dist + j) > (IO)0)
                        hit = j;
                if (hit < 0) break;
                src = hit;
            }
            if (src != l) { r = __ldg(rho + src); d = __ldg(dist + src); }
            Mat3 T;
            transition_matrix(herm_axpy((double)r, s_osc.vm, h0), T_SCALE * (double)d, T);
            if (first) {
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) M[b][a] = T[a][b]; // M holds columns: M[c][k]
                first = false;
            } else {
                times_right<3>(T, M);
            }
        }
        IO *o = probability + i * 9;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                // amplitude a -> b is A[b][a] = column a, component b
                const Cplx z = first ? Cplx{0.0, 0.0} : M[a][b];
                o[a * 3 + b] = (IO)fma(z.re, z.re, z.im * z.im);
            }
    }
}

// ---- neutrino decay (decay_flag == 1, prob3_decay.cuh): the same two entry points with the general-matrix layer ----
// One thread per event, state and per-event Hamiltonian in per-thread shared-memory columns like the standard kernels,
// plain loads (a ~50k-cycle event hides them).
constexpr int kDecayBlock = 128;
#ifndef PISAB_DECAY_EARTH_MIN_BLOCKS
#define PISAB_DECAY_EARTH_MIN_BLOCKS 3 // 168 registers; 2 / 3 / 4 blocks: 2.86 / 2.56 / 2.59 ms per 4e6 events (3x3 output, ordered)
#endif

template <bool FULL>
static size_t earth_decay_smem_bytes() {
    return (size_t)((FULL ? PropagatorSmem<3, 3>::kDoubles : PropagatorSmem<1, 2>::kDoubles) + H0DecaySmem::kDoubles) *
           kDecayBlock * sizeof(double);
}
template <typename IO, bool FULL>
__global__ void __launch_bounds__(kDecayBlock, PISAB_DECAY_EARTH_MIN_BLOCKS)
prob3_earth_decay_kernel(const __grid_constant__ OscTable osc, const __grid_constant__ DecayTable dec,
                         const __grid_constant__ EarthTable earth, int nubar, const int32_t *__restrict__ d_nubar,
                         int flav, const int32_t *__restrict__ d_flav, const IO *__restrict__ energy,
                         const IO *__restrict__ coszen, const int32_t *__restrict__ order, int64_t n,
                         IO *__restrict__ probability, IO *__restrict__ prob_e, IO *__restrict__ prob_mu) {
    constexpr int NR = FULL ? 3 : 1, NC = FULL ? 3 : 2;
    extern __shared__ __align__(16) double s_dyn_decay[];
    __shared__ EarthTable s_earth;
    double2(*s_state)[kDecayBlock] = reinterpret_cast<double2(*)[kDecayBlock]>(s_dyn_decay);
    double(*s_h0)[kDecayBlock] = reinterpret_cast<double(*)[kDecayBlock]>(s_dyn_decay + PropagatorSmem<NR, NC>::kDoubles * kDecayBlock);
    copy_earth(earth, &s_earth);
    const int tid = threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // `order` (optional): events grouped by crossed shells, so that a warp's lanes walk the same number of layers
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + tid; t < n; t += stride) {
        const int64_t i = order ? (int64_t)__ldg(order + t) : t;
        const double e = ld(energy, i), cz = ld(coszen, i);
        const int nb = d_nubar ? __ldg(d_nubar + i) : nubar;
        const int fl = d_flav ? __ldg(d_flav + i) : flav;
        const double inv_e = rcp_fast(e);
        H0DecaySmem h0{&s_h0[0][tid], kDecayBlock};
        h0.init(herm_axpy(nb > 0 ? inv_e : -inv_e, osc.hv[0], osc.lr), dec, nb, inv_e);
        PropagatorSmem<NR, NC> P{&s_state[0][tid], kDecayBlock};
        propagate_earth<NR, NC, false>(h0, osc, s_earth, cz, inv_e, nb, FULL ? 0 : fl, P);
        if (FULL) {
            IO *o = probability + i * 9;
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) o[a * 3 + b] = (IO)P.prob(b, a);
        } else {
            prob_e[i] = (IO)P.prob(0, 0);
            prob_mu[i] = (IO)P.prob(0, 1);
        }
    }
}

// propagate_array with decay_flag == 1 on explicit layers; same layer-cache rule as prob3_layers_kernel
template <typename IO>
__global__ void __launch_bounds__(kDecayBlock, PISAB_DECAY_EARTH_MIN_BLOCKS)
prob3_layers_decay_kernel(const __grid_constant__ OscTable osc, const __grid_constant__ DecayTable dec, int nubar,
                          const int32_t *__restrict__ d_nubar, const IO *__restrict__ energy,
                          const IO *__restrict__ densities, const IO *__restrict__ distances, int64_t n,
                          int n_layers, IO *__restrict__ probability) {
    const double T_SCALE = kTab[18];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double e = ld(energy, i);
        const int nb = d_nubar ? __ldg(d_nubar + i) : nubar;
        const double inv_e = 1.0 / e;
        H0Decay h0;
        h0.init(herm_axpy(nb > 0 ? inv_e : -inv_e, osc.hv[0], osc.lr), dec, nb, inv_e);
        const IO *rho = densities + i * n_layers;
        const IO *dist = distances + i * n_layers;
        Cplx M[3][3];
        bool first = true;
        for (int l = 0; l < n_layers; ++l) {
            IO d = __ldg(dist + l);
            if (!(d > (IO)0)) continue;
            IO r = __ldg(rho + l);
            int src = l;
            for (;;) {
                int hit = -1;
                const IO rs = __ldg(rho + src), ds = __ldg(dist + src);
                for (int j = 0; j < src; ++j)
                    if (fabs((double)(__ldg(rho + j) - rs)) < 1e-5 && fabs((double)(__ldg(dist + j) - ds)) < 1e-5 &&
                        __ldg(dist + j) > (IO)0)
                        hit = j;
                if (hit < 0) break;
                src = hit;
            }
            if (src != l) { r = __ldg(rho + src); d = __ldg(dist + src); }
            Mat3 T;
            h0.layer((double)r, osc.vm, T_SCALE * (double)d, T);
            if (first) {
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) M[b][a] = T[a][b];
                first = false;
            } else {
                times_right<3>(T, M);
            }
        }
        IO *o = probability + i * 9;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                const Cplx z = first ? Cplx{0.0, 0.0} : M[a][b];
                o[a * 3 + b] = (IO)fma(z.re, z.re, z.im * z.im);
            }
    }
}

// dynamic shared memory of reweight_hist_kernel (layout documented in the kernel)
template <typename IO>
static size_t fused_smem_bytes(int n_bins, bool std_matter, bool mp = false, bool decay = false) {
    // (FP32 mode: 9 float2 of state per thread = 9 doubles' worth; the Hamiltonian lives in registers)
    // (decay: the 18 doubles of the general per-event matrix take the place of the packed Hermitian one)
    const size_t h0_doubles = decay ? H0DecaySmem::kDoubles : (std_matter ? H0Smem<true>::kDoubles : H0Smem<false>::kDoubles);
    const size_t doubles = mp ? (size_t)PropagatorSmemF<1, 2>::kSlots * kBlock
                              : (size_t)(PropagatorSmem<1, 2>::kDoubles + h0_doubles) * kBlock;
    return WarpHist::smem_bytes(kBlock, n_bins) + doubles * sizeof(double) + (size_t)kBlock * (5 * sizeof(IO) + 4);
}

// Fused template evaluation: probabilities + reweighting + weighted histogram (w, w^2), for up to
// PISAB_MAX_BATCH flavour containers in ONE launch.
//
// Grid: one resident wave (persistent).  Every block walks its grid-stride share of EVERY container in
// turn (perfectly balanced whatever the container sizes), flushing its warp-private histograms to
// partials[container][block] between containers; a second tiny kernel sums the partials in block order.
//
// Memory latency: a thread spends ~20k cycles of FP64 work per event and then needs 44 bytes of the
// next one; with 4 warps per scheduler an exposed DRAM round trip (~1.5k cycles, twice: inputs at
// the top, flux/weight/bin at the bottom) costs ~15 % (ncu, round 1 capture d).  So every thread
// software-pipelines its own stream: at the top of an event it issues cp.async copies of (a) the
// flux / weight / bin of THIS event, needed at the bottom, and (b) energy / coszen of its NEXT event
// into its private shared-memory slots, and waits for them only after the propagation.  The index of
// the event after next (`order` indirection) rides in a register.
//
// Registers: the propagation state (row + two column vectors) and the per-event part of the
// Hamiltonian live in per-thread columns of shared memory, so the eigenvalue solve + matrix assembly
// fit 128 registers (2 x 256 threads per SM) without spills or loop-carried moves.
template <typename IO>
struct FusedContainer {
    const IO *energy, *coszen, *nu_flux, *weights_in;
    const int32_t *index, *order, *d_nubar, *d_flav;
    IO *weights_out, *prob_e, *prob_mu;
    int64_t n;
    double scale; // per-container factor folded into the weight (aeff.aeff: livetime * aeff_scale * norms)
    int32_t nubar, flav;
    int32_t flags; // PISAB_CONTAINER_*
    // PISAB_CONTAINER_FLUX_SYS: flux.barr_simple is evaluated in the kernel from these instead of reading nu_flux
    const double *flux_terms;
    const IO *nu_nom, *nubar_nom;
    const IO *astro; // optional additive per-event term (hist.py:141-145), non-PLAIN instantiations only
};
template <typename IO>
struct FusedBatch {
    int32_t n_containers;
    int32_t n_bins;
    BarrSys sys; // flux systematics of this template (FLUX instantiations)
    FusedContainer<IO> c[PISAB_MAX_BATCH];
};

// Body shared by the one-template kernel (oscillation table in the parameter constant bank) and the
// multi-template scan kernel (table of this block's template in shared memory): `rank` of `n_ranks`
// blocks cooperate on one template; partials: [container][n_ranks][2 n_bins] of that template.
// PLAIN: no per-event outputs and no per-event nubar / flav arrays (the fit-loop case): the null checks and
// the three optional stores disappear from the event loop.
// LARGE (n_bins > PISAB_DET_MAX_BINS): no private bins; every weight goes to the exact fixed-point accumulators in
// global memory (hist_device.cuh), `partials` then points at FixedAcc[container][2][n_bins] and `bounds` at the
// per-container weight bounds written by fused_bound_kernel.
// FLUX: containers flagged PISAB_CONTAINER_FLUX_SYS get their nu_flux from flux.barr_simple evaluated in registers
// (flux_device.cuh).  Its 64 bytes per event (cached terms + two nominal fluxes) do not fit the staging slots -- two
// blocks per SM already use all of the shared memory -- so they are prefetched into L2 at the top of the event and
// loaded after the propagation, when the registers are free; an L2 hit per ~16k-cycle event is hidden by the other
// warps.
// DECAY (decay_flag == 1): the general-matrix layers of prob3_decay.cuh with the state in registers; `dec` is the
// launch's DecayTable.  Everything around the propagation (staging, weights, histogram) is the common code.
template <typename IO, bool STD, bool PLAIN, bool MP = false, bool LARGE = false, bool FLUX = false, bool DECAY = false>
__device__ __forceinline__ void fused_template_body(const OscTable &osc, const EarthTable &s_earth,
                                                    const FusedBatch<IO> &batch, int ci_begin, int ci_end,
                                                    int rank, int n_ranks, double *__restrict__ partials,
                                                    double *s_hist,
                                                    const unsigned long long *__restrict__ bounds = nullptr,
                                                    int mode = 0, const DecayTable *dec = nullptr) {
    // dynamic shared memory: [histogram: warps x (2 n_bins + 32)] [per-thread state 9 x double2 x block]
    // [per-thread h0 (+ invariants + h0^2) x block] [flux 2 x block] [e, cz, w: block each] (IO) [bin: block]
    const int n_bins = batch.n_bins;
    // (LARGE: no bins, only the warps' [4][32] staging words of warp_fixed_add)
    double *s_dyn = s_hist + (LARGE ? (size_t)(kBlock / 32) * 128 : WarpHist::smem_bytes(kBlock, n_bins) / sizeof(double));
    unsigned long long *s_fixed = reinterpret_cast<unsigned long long *>(s_hist) + (threadIdx.x >> 5) * 128;
    double2(*s_state)[kBlock] = reinterpret_cast<double2(*)[kBlock]>(s_dyn);
    float2(*s_statef)[kBlock] = reinterpret_cast<float2(*)[kBlock]>(s_dyn); // FP32 mode: float2 state columns
    s_dyn += (MP ? PropagatorSmemF<1, 2>::kSlots : PropagatorSmem<1, 2>::kDoubles) * kBlock;
    double(*s_h0)[kBlock] = reinterpret_cast<double(*)[kBlock]>(s_dyn); // 16-byte aligned: see H0Smem
    if (!MP) s_dyn += (DECAY ? H0DecaySmem::kDoubles : H0Smem<STD>::kDoubles) * kBlock;   // (FP32 mode: the per-event Hamiltonian lives in registers)
    IO(*s_flux)[2] = reinterpret_cast<IO(*)[2]>(s_dyn);
    IO *s_e = &s_flux[kBlock][0], *s_cz = s_e + kBlock, *s_w = s_cz + kBlock;
    int32_t *s_bin = reinterpret_cast<int32_t *>(s_w + kBlock);
    WarpHist wh(s_hist, n_bins);
    const int tid = threadIdx.x;
    const int64_t stride = (int64_t)n_ranks * blockDim.x;
    // Which events a thread walks: t = first + k * stride.  Default: a block takes 256 consecutive events per round --
    // the events are sorted by crossed shells, so its warps cost the same and finish together (what a multi-wave grid
    // wants).  `interleave` (single-wave grids, i.e. analysis-size samples): consecutive 32-event chunks go round-robin
    // over the blocks instead, so that every SM gets the same mix of deep (expensive) and shallow (cheap) events; with
    // consecutive chunks the SM that holds the deepest events ran 4x longer than the one with the shallowest
    // (sm__cycles_active 13k .. 57k, profiles/r02_small_template.txt), and odd rounds walk the chunks BACKWARDS, so
    // that the warp that drew the deepest chunk of round 0 draws the shallowest of round 1 (the critical path of a
    // two-round template is one warp's two events).
    const bool interleave = mode & 1;
    const int first = interleave ? ((tid >> 5) * n_ranks + rank) * 32 + (tid & 31) : rank * (int)blockDim.x + tid;
    const int first_odd = (mode & 2) ? (int)stride - 32 - (first - (tid & 31)) + (tid & 31) : first;

    for (int ci = ci_begin; ci < ci_end; ++ci) {
        const FusedContainer<IO> &C = batch.c[ci];
        const IO *__restrict__ energy = C.energy, *__restrict__ coszen = C.coszen;
        const IO *__restrict__ nu_flux = C.nu_flux, *__restrict__ weights_in = C.weights_in;
        const int32_t *__restrict__ index = C.index, *__restrict__ order = C.order;
        const int64_t n = C.n;
        const bool fold_flux = FLUX && (C.flags & PISAB_CONTAINER_FLUX_SYS);
        FixedAcc *acc = nullptr;
        double sc1 = 1.0, sc2 = 1.0;
        if (LARGE) {
            acc = reinterpret_cast<FixedAcc *>(partials) + (size_t)ci * 2 * n_bins;
            const double bound = __longlong_as_double((long long)bounds[ci]);
            sc1 = fixed_scale(bound, (double)n);
            sc2 = fixed_scale(bound * bound, (double)n);
        } else {
            wh.clear();
        }
        // `order` (optional) lists the events grouped by number of crossed shells; -1 = no event
        // (n < 2^31 is checked by the host wrapper: 32-bit event indices save registers)
        auto event_of = [&](int64_t t) -> int { return t < n ? (order ? __ldg(order + t) : (int)t) : -1; };
        int i_cur = event_of(first), i_next = event_of(stride + first_odd);
        if (i_cur >= 0) { s_e[tid] = __ldg(energy + i_cur); s_cz[tid] = __ldg(coszen + i_cur); }
        // block-uniform trip count (rounds of `stride` events), so that the warp-collective histogram step is converged
        int odd = 0;
        for (int64_t round = 0; round < n; round += stride, odd ^= 1) {
            const int i_nn = event_of(round + 2 * stride + (odd ? first_odd : first));
            double w = 0.0;
            int bin = -1;
            if (i_cur >= 0) {
                const int64_t i = i_cur;
                const double e = (double)s_e[tid], cz = (double)s_cz[tid];
                if (fold_flux) {
                    prefetch_l2(C.flux_terms + 4 * i);
                    prefetch_l2(C.nu_nom + 2 * i);
                    prefetch_l2(C.nubar_nom + 2 * i);
                } else {
                    cp_async<2 * sizeof(IO)>(&s_flux[tid][0], nu_flux + 2 * i);
                }
                cp_async<sizeof(IO)>(&s_w[tid], weights_in + i);
                cp_async<4>(&s_bin[tid], index + i);
                if (i_next >= 0) {
                    cp_async<sizeof(IO)>(&s_e[tid], energy + i_next);
                    cp_async<sizeof(IO)>(&s_cz[tid], coszen + i_next);
                }
                cp_async_commit();
                const int nb = (!PLAIN && C.d_nubar) ? __ldg(C.d_nubar + i) : C.nubar;
                const int fl = (!PLAIN && C.d_flav) ? __ldg(C.d_flav + i) : C.flav;
                const double inv_e = rcp_fast(e);
                const Herm3 hh = herm_axpy(nb > 0 ? inv_e : -inv_e, osc.hv[0], osc.lr); // hv[1] = -hv[0]
                double pe, pmu;
                if constexpr (DECAY) {
                    H0DecaySmem h0{&s_h0[0][tid], kBlock};
                    h0.init(hh, *dec, nb, inv_e);
                    PropagatorSmem<1, 2> P{&s_state[0][tid], kBlock};
                    propagate_earth<1, 2, false>(h0, osc, s_earth, cz, inv_e, nb, fl, P);
                    pe = P.prob(0, 0);
                    pmu = P.prob(0, 1);
                } else if constexpr (MP) {
                    H0MP<STD> h0;
                    h0.init(hh);
                    PropagatorSmemF<1, 2> P{&s_statef[0][tid], kBlock};
                    propagate_earth<1, 2, STD>(h0, osc, s_earth, cz, inv_e, nb, fl, P);
                    pe = P.prob(0, 0);
                    pmu = P.prob(0, 1);
                } else {
                    H0Smem<STD> h0{&s_h0[0][tid], kBlock};
                    h0.store(hh);
                    if (STD) h0.set_poly(hh);
                    PropagatorSmem<1, 2> P{&s_state[0][tid], kBlock};
                    propagate_earth<1, 2, STD>(h0, osc, s_earth, cz, inv_e, nb, fl, P);
                    pe = P.prob(0, 0);
                    pmu = P.prob(0, 1);
                }
                cp_async_wait_all();
                // prob3.py:622: weights *= (flux_e * prob_e) + (flux_mu * prob_mu)
                double fe, fm;
                if (fold_flux) {
                    barr_apply_event<IO>(batch.sys, C.flux_terms, C.nu_nom, C.nubar_nom, nb, i, fe, fm);
                    fe = (double)(IO)fe; // rounded to the storage type like the nu_flux array it replaces
                    fm = (double)(IO)fm;
                } else {
                    fe = (double)s_flux[tid][0];
                    fm = (double)s_flux[tid][1];
                }
                w = (double)s_w[tid] * (fe * pe + fm * pmu) * C.scale;
                bin = s_bin[tid];
                if (!PLAIN) {
                    if (C.astro) w += (double)__ldg(C.astro + i);
                    if (C.weights_out) C.weights_out[i] = (IO)w;
                    if (C.prob_e) C.prob_e[i] = (IO)pe;
                    if (C.prob_mu) C.prob_mu[i] = (IO)pmu;
                }
            }
            if (LARGE) {
                warp_fixed_add(s_fixed, acc, n_bins, bin, w, sc1, sc2);
            } else {
                wh.add(bin, w);
            }
            i_cur = i_next;
            i_next = i_nn;
        }
        if (!LARGE) {
            wh.flush(partials + ((size_t)ci * n_ranks + rank) * 2 * n_bins);
            __syncthreads(); // the next container clears the bins
        }
    }
}

// Weight bound per container for the fixed-point accumulators: max over events of |w_in| (|flux_e| + |flux_mu|) |scale|
// >= |w| because both probabilities are <= 1.  24 B/event, HBM-bound (4 % of the fused kernel's time).
template <typename IO>
__global__ void __launch_bounds__(256)
fused_bound_kernel(const __grid_constant__ FusedBatch<IO> batch, int blocks_per_container,
                   unsigned long long *__restrict__ bounds) {
    __shared__ double s[256];
    const int ci = blockIdx.x / blocks_per_container, r = blockIdx.x - ci * blocks_per_container;
    const FusedContainer<IO> &C = batch.c[ci];
    double m = 0.0;
    const int64_t stride = (int64_t)blocks_per_container * blockDim.x;
    for (int64_t i = (int64_t)r * blockDim.x + threadIdx.x; i < C.n; i += stride)
        m = fmax(m, fabs((double)__ldg(C.weights_in + i)) *
                        (fabs((double)__ldg(C.nu_flux + 2 * i)) + fabs((double)__ldg(C.nu_flux + 2 * i + 1))));
    s[threadIdx.x] = m * fabs(C.scale);
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) s[threadIdx.x] = fmax(s[threadIdx.x], s[threadIdx.x + off]);
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicMax(bounds + ci, (unsigned long long)__double_as_longlong(s[0]));
}

template <typename IO>
__global__ void __launch_bounds__(256)
fused_fixed_finish_kernel(const __grid_constant__ FusedBatch<IO> batch, const FixedAcc *__restrict__ acc,
                          const unsigned long long *__restrict__ bounds, double *__restrict__ out) {
    const int n_bins = batch.n_bins;
    const int v = blockIdx.x * blockDim.x + threadIdx.x; // (container, plane, bin)
    if (v >= batch.n_containers * 2 * n_bins) return;
    const int ci = v / (2 * n_bins), plane = (v - ci * 2 * n_bins) / n_bins;
    const double bound = __longlong_as_double((long long)bounds[ci]);
    const double sc = fixed_scale(plane ? bound * bound : bound, (double)batch.c[ci].n);
    out[v] = fixed_value(acc[v], sc);
}

template <typename IO, bool STD, bool PLAIN, bool MP = false, bool LARGE = false, bool FLUX = false>
__global__ void __launch_bounds__(kBlock, MP ? PISAB_MP_MIN_BLOCKS : PISAB_MIN_BLOCKS)
reweight_hist_kernel(const __grid_constant__ OscTable osc, const __grid_constant__ EarthTable earth,
                     const __grid_constant__ FusedBatch<IO> batch, int ranks, double *__restrict__ partials,
                     const unsigned long long *__restrict__ bounds, int interleave, const __grid_constant__ FusedEpi epi) {
    extern __shared__ __align__(16) double s_hist[];
    __shared__ EarthTable s_earth;
    // the oscillation table is read straight from the kernel-parameter constant bank (fixed offsets: a
    // DFMA takes such an operand directly); the Earth table is indexed per lane and goes to shared memory
    copy_earth(earth, &s_earth);
    // block -> (container, rank): `ranks` blocks share one container (grid = n_containers * ranks).  A block
    // never walks more than one container, so a template over analysis-size containers (1e4 events each)
    // costs one or two event latencies instead of one per container.
    const int ci = blockIdx.x / ranks, rank = blockIdx.x - ci * ranks;
    fused_template_body<IO, STD, PLAIN, MP, LARGE, FLUX>(osc, s_earth, batch, ci, ci + 1, rank, ranks, partials, s_hist, bounds,
                                                         interleave);
    // one hypothesis = one launch: the last block of every container reduces it, the last container sums and scores
    if (!LARGE && epi.out) fused_epilogue(epi, partials, ci, ranks, batch.n_containers, batch.n_bins, s_hist);
}

// The template kernel of the decay branch: same grid / partials / epilogue contract as reweight_hist_kernel (so the
// host code after the launch is shared), per-event outputs and in-kernel flux.barr_simple included; one instantiation
// per storage type; state and per-event Hamiltonian in the per-thread shared-memory columns of the standard kernel.
#ifndef PISAB_DECAY_MIN_BLOCKS
#define PISAB_DECAY_MIN_BLOCKS 2 // 128 registers (some spills) beat one block at 255: 15.6 vs 16.9 ms per 4.8e7 events
#endif
template <typename IO>
__global__ void __launch_bounds__(kBlock, PISAB_DECAY_MIN_BLOCKS)
reweight_hist_decay_kernel(const __grid_constant__ OscTable osc, const __grid_constant__ DecayTable dec,
                           const __grid_constant__ EarthTable earth, const __grid_constant__ FusedBatch<IO> batch,
                           int ranks, double *__restrict__ partials, const __grid_constant__ FusedEpi epi) {
    extern __shared__ __align__(16) double s_hist[];
    __shared__ EarthTable s_earth;
    copy_earth(earth, &s_earth);
    const int ci = blockIdx.x / ranks, rank = blockIdx.x - ci * ranks;
    fused_template_body<IO, false, false, false, false, true, true>(osc, s_earth, batch, ci, ci + 1, rank, ranks, partials,
                                                                    s_hist, nullptr, 0, &dec);
    if (epi.out) fused_epilogue(epi, partials, ci, ranks, batch.n_containers, batch.n_bins, s_hist);
}

// ---- FP32 mode, TWO events per thread (prob3_mp.cuh: the float part of a pair runs in the two lanes of the packed
// FP32 instructions, halving its issue slots per event; FP64 eigenvalues / phases / geometry stay per event) ----------
// Requirements, guaranteed by the host (PISAB_CONTAINER_PAIR_ALIGNED): float storage, an even number of events,
// events 2k and 2k+1 cross the same Earth shells, no `order` indirection, no per-event outputs.
// Block of 128 threads, four blocks per SM (128 registers): the same 16 warps per SM as 2 x 256, but finer-grained
// (measured +6 %, profiles/r02_pair_kernel_variants.txt); three blocks of 256 at 85 registers spill (-25 %).
// The block size is a template parameter: multi-wave grids of the standard-matter instantiation run 2.2 % faster
// still with 64 threads x 8 blocks (6.79 instead of 6.94 ms per 9.6e7 events; NSI: level), while analysis-size
// single-wave templates want the 128-thread blocks (64: 66 instead of 35 us per 1.2e5-event template, 32 x 16: 101 us;
// profiles/r02_pair_kernel_variants.txt) -- the host picks (reweight_hist_batch_impl).
#ifndef PISAB_PAIR_BLOCK
#define PISAB_PAIR_BLOCK 128
#endif
constexpr int kPairBlock = PISAB_PAIR_BLOCK;
constexpr int kPairBlockSmall = 64; // multi-wave grids, standard matter
static size_t fused_pair_smem_bytes(int n_bins, int block = kPairBlock) {
    return WarpHist::smem_bytes(block, n_bins) +
           (size_t)block * (PropagatorSmemP<1, 2>::kSlots * sizeof(float4) + H0MP2<true>::kSlots * sizeof(float2) + 48) +
           (size_t)(block / 32) * 32 * sizeof(double); // second staging row of WarpHist::add2
}

template <bool STD, bool FLUX, bool MIX, int kPairBlock>
__device__ __forceinline__ void fused_pair_body(const OscTable &osc, const EarthTable &s_earth,
                                                const FusedBatch<float> &batch, int ci, int rank, int n_ranks,
                                                double *__restrict__ partials, double *s_hist, int mode) {
    // dynamic shared memory: [histogram] [state: 9 float4 x block] [flux float4] [e, cz, w: float2 each] [bin int2]
    const int n_bins = batch.n_bins;
    unsigned char *s_dyn = reinterpret_cast<unsigned char *>(s_hist) + WarpHist::smem_bytes(kPairBlock, n_bins);
    float4(*s_state)[kPairBlock] = reinterpret_cast<float4(*)[kPairBlock]>(s_dyn);
    s_dyn += (size_t)PropagatorSmemP<1, 2>::kSlots * kPairBlock * sizeof(float4);
    float2(*s_base)[kPairBlock] = reinterpret_cast<float2(*)[kPairBlock]>(s_dyn);     // packed per-pair invariants (H0MP2)
    s_dyn += (size_t)H0MP2<true>::kSlots * kPairBlock * sizeof(float2);
    float4 *s_flux = reinterpret_cast<float4 *>(s_dyn);
    float2 *s_e = reinterpret_cast<float2 *>(s_flux + kPairBlock), *s_cz = s_e + kPairBlock, *s_w = s_cz + kPairBlock;
    int2 *s_bin = reinterpret_cast<int2 *>(s_w + kPairBlock);
    double *s_stage1 = reinterpret_cast<double *>(s_bin + kPairBlock) + (threadIdx.x >> 5) * 32;
    WarpHist wh(s_hist, n_bins);
    const int tid = threadIdx.x;
    const int64_t stride = (int64_t)n_ranks * blockDim.x;
    // MIX (single-wave grids): chunks interleaved over the blocks, odd rounds backwards -- see fused_template_body; a
    // template parameter here because this kernel sits at the edge of its register budget (the run-time form cost
    // 1.6 % of the multi-wave throughput)
    const bool interleave = MIX && (mode & 1);
    const int first = interleave ? ((tid >> 5) * n_ranks + rank) * 32 + (tid & 31) : rank * (int)blockDim.x + tid;
    const int first_odd = (MIX && (mode & 2)) ? (int)stride - 32 - (first - (tid & 31)) + (tid & 31) : first;
    const FusedContainer<float> &C = batch.c[ci];
    const float2 *__restrict__ energy = reinterpret_cast<const float2 *>(C.energy);
    const float2 *__restrict__ coszen = reinterpret_cast<const float2 *>(C.coszen);
    const float4 *__restrict__ nu_flux = reinterpret_cast<const float4 *>(C.nu_flux);
    const float2 *__restrict__ weights_in = reinterpret_cast<const float2 *>(C.weights_in);
    const int2 *__restrict__ index = reinterpret_cast<const int2 *>(C.index);
    const int64_t n_pairs = C.n >> 1;
    const bool fold_flux = FLUX && (C.flags & PISAB_CONTAINER_FLUX_SYS);
    wh.clear();
    auto pair_of = [&](int64_t t) -> int { return t < n_pairs ? (int)t : -1; };
    int p_cur = pair_of(first), p_next = pair_of(stride + first_odd);
    if (p_cur >= 0) { s_e[tid] = __ldg(energy + p_cur); s_cz[tid] = __ldg(coszen + p_cur); }
    int odd = 0;
    for (int64_t round = 0; round < n_pairs; round += stride, odd ^= 1) {
        const int p_nn = pair_of(round + 2 * stride + (odd ? first_odd : first));
        double w0 = 0.0, w1 = 0.0;
        int bin0 = -1, bin1 = -1;
        if (p_cur >= 0) {
            const float2 e2 = s_e[tid], c2 = s_cz[tid];
            if (fold_flux) {
                prefetch_l2(C.flux_terms + 8 * (int64_t)p_cur);
                prefetch_l2(C.flux_terms + 8 * (int64_t)p_cur + 4);
                prefetch_l2(C.nu_nom + 4 * (int64_t)p_cur);
                prefetch_l2(C.nubar_nom + 4 * (int64_t)p_cur);
            } else {
                cp_async<16>(&s_flux[tid], nu_flux + p_cur);
            }
            cp_async<8>(&s_w[tid], weights_in + p_cur);
            cp_async<8>(&s_bin[tid], index + p_cur);
            if (p_next >= 0) {
                cp_async<8>(&s_e[tid], energy + p_next);
                cp_async<8>(&s_cz[tid], coszen + p_next);
            }
            cp_async_commit();
            const double cz[2] = {(double)c2.x, (double)c2.y};
            const double inv_e[2] = {rcp_fast((double)e2.x), rcp_fast((double)e2.y)};
            const double sg = C.nubar > 0 ? 1.0 : -1.0;
            H0MP2<STD> h0;
            h0.col = &s_base[0][tid];
            h0.pitch = kPairBlock;
            h0.init(herm_axpy(sg * inv_e[0], osc.hv[0], osc.lr), herm_axpy(sg * inv_e[1], osc.hv[0], osc.lr)); // hv[1] = -hv[0]
            PropagatorSmemP<1, 2> P{&s_state[0][tid], kPairBlock};
            bool mismatch;
            propagate_earth_pair<1, 2, STD>(h0, osc, s_earth, cz, inv_e, C.nubar, C.flav, P, mismatch);
            const f2 pe = P.prob_r(0, 0), pmu = P.prob_r(0, 1);
            cp_async_wait_all();
            float4 fl; // (flux_e, flux_mu) of event 0, then of event 1
            if (fold_flux) {
                double a0, a1, b0, b1;
                barr_apply_event<float>(batch.sys, C.flux_terms, C.nu_nom, C.nubar_nom, C.nubar, 2 * (int64_t)p_cur, a0, a1);
                barr_apply_event<float>(batch.sys, C.flux_terms, C.nu_nom, C.nubar_nom, C.nubar, 2 * (int64_t)p_cur + 1, b0, b1);
                fl = make_float4((float)a0, (float)a1, (float)b0, (float)b1);
            } else {
                fl = s_flux[tid];
            }
            const float2 ww = s_w[tid];
            const int2 bb = s_bin[tid];
            // prob3.py:622: weights *= (flux_e * prob_e) + (flux_mu * prob_mu)
            w0 = (double)ww.x * ((double)fl.x * (double)pe.x + (double)fl.y * (double)pmu.x) * C.scale;
            w1 = (double)ww.y * ((double)fl.z * (double)pe.y + (double)fl.w * (double)pmu.y) * C.scale;
            if (mismatch) w0 = w1 = __longlong_as_double(0x7ff8000000000000LL); // the host broke the pairing contract
            bin0 = bb.x;
            bin1 = bb.y;
        }
        wh.add2(bin0, w0, bin1, w1, s_stage1);
        p_cur = p_next;
        p_next = p_nn;
    }
    wh.flush(partials + ((size_t)ci * n_ranks + rank) * 2 * n_bins);
}

template <bool STD, bool FLUX = false, bool MIX = false, int BLOCK = kPairBlock>
__global__ void __launch_bounds__(BLOCK, 512 / BLOCK)
reweight_hist_pair_kernel(const __grid_constant__ OscTable osc, const __grid_constant__ EarthTable earth,
                          const __grid_constant__ FusedBatch<float> batch, int ranks, double *__restrict__ partials,
                          const unsigned long long *__restrict__ /* bounds: same signature as reweight_hist_kernel */,
                          int interleave, const __grid_constant__ FusedEpi epi) {
    extern __shared__ __align__(16) double s_hist[];
    __shared__ EarthTable s_earth;
    copy_earth(earth, &s_earth);
    const int ci = blockIdx.x / ranks, rank = blockIdx.x - ci * ranks;
    fused_pair_body<STD, FLUX, MIX, BLOCK>(osc, s_earth, batch, ci, rank, ranks, partials, s_hist, interleave);
    if (MIX && epi.out) fused_epilogue(epi, partials, ci, ranks, batch.n_containers, batch.n_bins, s_hist);
}

// Parameter scan (BASELINE configs[4]): P templates in ONE launch.  Block b serves (template, container, rank)
// = (b / (C R), (b / R) % C, b % R); its oscillation table comes from a device array (P tables do not fit the
// parameter space) and is staged in shared memory.  For the event samples of a real analysis (1e5 .. 1e6
// events) one template does not fill the GPU and a per-template launch is bound by launch latency plus one or
// two event latencies (~45 us); batching the hypotheses restores full occupancy.  Per-event outputs are not written (they would race between templates).
template <typename IO, bool STD, bool MP = false>
__global__ void __launch_bounds__(kBlock, MP ? PISAB_MP_MIN_BLOCKS : PISAB_MIN_BLOCKS)
reweight_hist_scan_kernel(const OscTable *__restrict__ tables, const __grid_constant__ EarthTable earth,
                          const __grid_constant__ FusedBatch<IO> batch, int ranks_per_template,
                          double *__restrict__ partials) {
    extern __shared__ __align__(16) double s_hist[];
    __shared__ EarthTable s_earth;
    __shared__ OscTable s_osc;
    // block -> (template, container, rank): as in reweight_hist_kernel a block serves one container
    const int per_template = batch.n_containers * ranks_per_template;
    const int tmpl = blockIdx.x / per_template, rem = blockIdx.x - tmpl * per_template;
    const int ci = rem / ranks_per_template, rank = rem - ci * ranks_per_template;
    {
        const double *src = reinterpret_cast<const double *>(tables + tmpl);
        double *dst = reinterpret_cast<double *>(&s_osc);
        for (int i = threadIdx.x; i < (int)(sizeof(OscTable) / 8); i += blockDim.x) dst[i] = __ldg(src + i);
    }
    copy_earth(earth, &s_earth); // ends with __syncthreads()
    double *mine = partials + (size_t)tmpl * batch.n_containers * ranks_per_template * 2 * batch.n_bins;
    fused_template_body<IO, STD, true, MP>(s_osc, s_earth, batch, ci, ci + 1, rank, ranks_per_template, mine, s_hist);
}

// The scan with neutrino decay: every template carries a DecayTable next to its OscTable (a template without decay
// has a zero table and goes through the same general-matrix layers: equal to the standard kernels within 2e-12).
template <typename IO>
__global__ void __launch_bounds__(kBlock, PISAB_DECAY_MIN_BLOCKS)
reweight_hist_scan_decay_kernel(const OscTable *__restrict__ tables, const DecayTable *__restrict__ dtables,
                                const __grid_constant__ EarthTable earth, const __grid_constant__ FusedBatch<IO> batch,
                                int ranks_per_template, double *__restrict__ partials) {
    extern __shared__ __align__(16) double s_hist[];
    __shared__ EarthTable s_earth;
    __shared__ OscTable s_osc;
    __shared__ DecayTable s_dec;
    const int per_template = batch.n_containers * ranks_per_template;
    const int tmpl = blockIdx.x / per_template, rem = blockIdx.x - tmpl * per_template;
    const int ci = rem / ranks_per_template, rank = rem - ci * ranks_per_template;
    {
        const double *src = reinterpret_cast<const double *>(tables + tmpl);
        double *dst = reinterpret_cast<double *>(&s_osc);
        for (int i = threadIdx.x; i < (int)(sizeof(OscTable) / 8); i += blockDim.x) dst[i] = __ldg(src + i);
        const double *srcd = reinterpret_cast<const double *>(dtables + tmpl);
        double *dstd = reinterpret_cast<double *>(&s_dec);
        for (int i = threadIdx.x; i < (int)(sizeof(DecayTable) / 8); i += blockDim.x) dstd[i] = __ldg(srcd + i);
    }
    copy_earth(earth, &s_earth); // ends with __syncthreads()
    double *mine = partials + (size_t)tmpl * batch.n_containers * ranks_per_template * 2 * batch.n_bins;
    fused_template_body<IO, false, true, false, false, false, true>(s_osc, s_earth, batch, ci, ci + 1, rank, ranks_per_template,
                                                                    mine, s_hist, nullptr, 0, &s_dec);
}

} // namespace pisab

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
using namespace pisab;

// kernel selection: the mixed-precision instantiations exist for float storage only
template <typename IO, bool FULL, bool STD>
static auto earth_kernel(bool mp) {
    if constexpr (sizeof(IO) == 4) {
        if (mp) return prob3_earth_kernel<IO, FULL, STD, true>;
    }
    return prob3_earth_kernel<IO, FULL, STD, false>;
}
template <typename IO, bool STD, bool PLAIN>
static auto fused_kernel(bool mp) {
    if constexpr (sizeof(IO) == 4) {
        if (mp) return reweight_hist_kernel<IO, STD, PLAIN, true, false>;
    }
    return reweight_hist_kernel<IO, STD, PLAIN, false, false>;
}
template <typename IO, bool STD>
static auto fused_kernel_flux(bool mp) { // flux.barr_simple folded in: fit-loop (PLAIN) form, <= PISAB_DET_MAX_BINS bins
    if constexpr (sizeof(IO) == 4) {
        if (mp) return reweight_hist_kernel<IO, STD, true, true, false, true>;
    }
    return reweight_hist_kernel<IO, STD, true, false, false, true>;
}
template <typename IO, bool STD>
static auto fused_kernel_large(bool mp) { // > PISAB_DET_MAX_BINS: fixed-point accumulators, fit-loop (PLAIN) form only
    if constexpr (sizeof(IO) == 4) {
        if (mp) return reweight_hist_kernel<IO, STD, true, true, true>;
    }
    return reweight_hist_kernel<IO, STD, true, false, true>;
}
template <typename IO, bool STD>
static auto scan_kernel(bool mp) {
    if constexpr (sizeof(IO) == 4) {
        if (mp) return reweight_hist_scan_kernel<IO, STD, true>;
    }
    return reweight_hist_scan_kernel<IO, STD, false>;
}

static int grid_for(int64_t n, int blocks_per_sm) {
    const int sms = sm_count();
    int64_t want = (n + kBlock - 1) / kBlock;
    int64_t cap = (int64_t)(sms > 0 ? sms : 148) * blocks_per_sm;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

// Grid sizing.  PISAB_WAVES: whole waves of resident blocks for kernels whose blocks all do the same work (scan,
// propagate); PISAB_SPREAD_WAVES: cap for the batched template kernel, whose blocks serve one container each.
#ifndef PISAB_WAVES
#define PISAB_WAVES 4
#endif
#ifndef PISAB_SPREAD_WAVES
#define PISAB_SPREAD_WAVES 32
#endif

// PISAB_WAVES static waves of blocks instead of one persistent wave: a block's share of the deep-core events
// varies, and with one wave nothing refills an SM whose blocks finish early (measured tail: ~3-6 %); the
// hardware scheduler balances several smaller waves, and the block -> events map stays static, so histograms
// remain bit-reproducible.  Whole waves only, and every thread keeps >= 8 events.
static int waved_grid(int resident, int64_t n_max) {
    const int64_t by_work = (n_max + (int64_t)kBlock * 8 - 1) / ((int64_t)kBlock * 8);
    int64_t waves = by_work / (resident > 0 ? resident : 1);
    if (waves > PISAB_WAVES) waves = PISAB_WAVES;
    return waves > 1 ? (int)(resident * waves) : resident;
}

template <typename K>
static int resident_grid(K kernel, int64_t n, size_t smem) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kBlock, smem) != cudaSuccess || occ < 1) occ = 2;
    if (occ > 8) occ = 8; // workspace bound (pisab_hist_workspace_bytes)
    return grid_for(n, occ);
}

template <typename IO>
static int propagate_earth_impl(const pisab_osc_consts_t *consts, const pisab_earth_t *earth,
                                int32_t nubar, const int32_t *d_nubar, int32_t flav,
                                const int32_t *d_flav, const IO *d_energy, const IO *d_coszen,
                                const int32_t *d_order, int64_t n, IO *d_probability, IO *d_prob_e,
                                IO *d_prob_mu, void *stream) {
    if (n < 0 || (n > 0 && (!d_energy || !d_coszen))) { set_error("bad event arrays"); return PISAB_ERR_ARG; }
    if (n > 2147483647LL) { set_error("at most 2^31-1 events per call (32-bit event indices)"); return PISAB_ERR_ARG; }
    if (!d_nubar && nubar != 1 && nubar != -1) { set_error("nubar must be +1 or -1"); return PISAB_ERR_ARG; }
    if ((d_prob_e == nullptr) != (d_prob_mu == nullptr)) { set_error("prob_e and prob_mu go together"); return PISAB_ERR_ARG; }
    if (d_prob_e && !d_flav && (flav < 0 || flav > 2)) { set_error("flav must be 0, 1 or 2"); return PISAB_ERR_ARG; }
    OscTable ot;
    EarthTable et;
    int rc = build_osc_table(consts, &ot);
    if (rc) return rc;
    rc = build_earth_table(earth, &et);
    if (rc) return rc;
    if (n == 0) return PISAB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (consts->decay_flag == 1) {
        // neutrino decay (numba_osc_kernels.py:445-451): general-matrix layers, FP64 arithmetic whatever the storage
        DecayTable dt;
        rc = build_decay_table(consts, &dt);
        if (rc) return rc;
        if (d_prob_e && d_probability && d_flav) { set_error("per-event flav with full probability output: call fill_probs per flavour"); return PISAB_ERR_ARG; }
        const int grid = grid_for(n, 2 * PISAB_DECAY_EARTH_MIN_BLOCKS) * (kBlock / kDecayBlock);
        LaunchTimer t(s);
        if (d_probability) {
            auto kernel = prob3_earth_decay_kernel<IO, true>;
            const size_t smem = earth_decay_smem_bytes<true>();
            PISAB_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kernel<<<grid, kDecayBlock, smem, s>>>(ot, dt, et, nubar, d_nubar, flav, d_flav, d_energy, d_coszen, d_order, n,
                                                   d_probability, nullptr, nullptr);
            note_launch();
        }
        if (d_prob_e) {
            auto kernel = prob3_earth_decay_kernel<IO, false>;
            const size_t smem = earth_decay_smem_bytes<false>();
            PISAB_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kernel<<<grid, kDecayBlock, smem, s>>>(ot, dt, et, nubar, d_nubar, flav, d_flav, d_energy, d_coszen, d_order, n,
                                                   nullptr, d_prob_e, d_prob_mu);
            note_launch();
        }
        PISAB_CUDA_CHECK(cudaGetLastError());
        return PISAB_OK;
    }
    const bool std_matter = ot.std_matter != 0.0;
    const bool mp = sizeof(IO) == 4 && f32_math_mixed();
    if (d_probability) {
        // the 3x3 state (74 KB per block) leaves no room for the cached H0^2 of the standard-matter path
        auto kernel = earth_kernel<IO, true, false>(mp);
        const size_t smem = earth_smem_bytes<true>(sizeof(IO), false, mp);
        PISAB_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LaunchTimer t(s);
        kernel<<<waved_grid(resident_grid(kernel, n, smem), n), kBlock, smem, s>>>(
            ot, et, nubar, d_nubar, flav, d_flav, d_energy, d_coszen, d_order, n, d_probability, nullptr, nullptr);
        note_launch();
    }
    if (d_prob_e && d_probability) {
        // fill_probs from the full matrix keeps the two outputs bit-consistent
        int r1 = sizeof(IO) == 8
                     ? pisab_fill_probs_f64((const double *)d_probability, 0, flav, n, (double *)d_prob_e, stream)
                     : pisab_fill_probs_f32((const float *)d_probability, 0, flav, n, (float *)d_prob_e, stream);
        if (r1) return r1;
        if (d_flav) { set_error("per-event flav with full probability output: call fill_probs per flavour"); return PISAB_ERR_ARG; }
        r1 = sizeof(IO) == 8
                 ? pisab_fill_probs_f64((const double *)d_probability, 1, flav, n, (double *)d_prob_mu, stream)
                 : pisab_fill_probs_f32((const float *)d_probability, 1, flav, n, (float *)d_prob_mu, stream);
        if (r1) return r1;
    } else if (d_prob_e) {
        auto kernel = std_matter ? earth_kernel<IO, false, true>(mp) : earth_kernel<IO, false, false>(mp);
        const size_t smem = earth_smem_bytes<false>(sizeof(IO), std_matter, mp);
        PISAB_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LaunchTimer t(s);
        kernel<<<waved_grid(resident_grid(kernel, n, smem), n), kBlock, smem, s>>>(
            ot, et, nubar, d_nubar, flav, d_flav, d_energy, d_coszen, d_order, n, nullptr, d_prob_e, d_prob_mu);
        note_launch();
    }
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}

template <typename IO>
static int propagate_layers_impl(const pisab_osc_consts_t *consts, int32_t nubar,
                                 const int32_t *d_nubar, const IO *d_energy, const IO *d_densities,
                                 const IO *d_distances, int64_t n, int32_t n_layers,
                                 IO *d_probability, void *stream) {
    if (n < 0 || (n > 0 && (!d_energy || !d_densities || !d_distances || !d_probability))) {
        set_error("bad event arrays");
        return PISAB_ERR_ARG;
    }
    // The reference caches at most 120 layer matrices (numba_osc_kernels.py:173-177,227) but loops over the arrays' width:
    // PREM_59layer arrives 122 wide (2 * 61 shells, at most 118 active) and works there.  The kernels here keep no
    // per-layer array, so the width is only bounded by the shell table.
    if (n_layers < 1 || n_layers > 2 * PISAB_MAX_RADII) {
        set_error("n_layers = %d outside [1, %d]", n_layers, 2 * PISAB_MAX_RADII);
        return PISAB_ERR_ARG;
    }
    if (!d_nubar && nubar != 1 && nubar != -1) { set_error("nubar must be +1 or -1"); return PISAB_ERR_ARG; }
    OscTable ot;
    int rc = build_osc_table(consts, &ot);
    if (rc) return rc;
    if (n == 0) return PISAB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (consts->decay_flag == 1) {
        DecayTable dt;
        rc = build_decay_table(consts, &dt);
        if (rc) return rc;
        LaunchTimer t(s);
        prob3_layers_decay_kernel<IO><<<grid_for(n, 2 * PISAB_DECAY_EARTH_MIN_BLOCKS) * (kBlock / kDecayBlock), kDecayBlock, 0, s>>>(
            ot, dt, nubar, d_nubar, d_energy, d_densities, d_distances, n, n_layers, d_probability);
        note_launch();
    } else {
        LaunchTimer t(s);
        prob3_layers_kernel<IO><<<grid_for(n, 4), kBlock, 0, s>>>(ot, nubar, d_nubar, d_energy, d_densities,
                                                                  d_distances, n, n_layers, d_probability);
        note_launch();
    }
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}

// Arrival words of the epilogue (hist_device.cuh: fused_epilogue / hist_reduce_chi2_kernel): 64 zero-initialised words
// per (device, stream), owned by the library.  The kernels reset what they counted before they exit, so consecutive
// launches on a stream reuse them; launches on different streams get different words.  nullptr (table full or
// allocation failed) sends the caller to the two-launch form.
unsigned *pisab::epilogue_counter(cudaStream_t stream) {
    struct Slot { int dev; cudaStream_t stream; unsigned *words; };
    static Slot slots[64];
    static int n_slots = 0;
    static std::mutex lock;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> guard(lock);
    for (int i = 0; i < n_slots; ++i)
        if (slots[i].dev == dev && slots[i].stream == stream) return slots[i].words;
    if (n_slots == 64) return nullptr;
    unsigned *p = nullptr;
    if (cudaMalloc(&p, 256) != cudaSuccess) return nullptr;
    if (cudaMemset(p, 0, 256) != cudaSuccess) { cudaFree(p); return nullptr; }
    slots[n_slots++] = Slot{dev, stream, p};
    return p;
}

// optional fit-loop epilogue of a batched template (pisab_reweight_hist_chi2_*)
struct Chi2Epilogue {
    const double *d_bin_scales, *d_observed;
    double *d_total, *d_chi2;
};

template <typename IO>
static int reweight_hist_batch_impl(const pisab_osc_consts_t *consts, const pisab_earth_t *earth,
                                    const FusedBatch<IO> &batch, int64_t n_max, double *d_batch_out,
                                    double *d_hist, double *d_hist_w2, void *d_workspace,
                                    int64_t workspace_bytes, void *stream, const Chi2Epilogue *epi = nullptr,
                                    bool has_sys = false) {
    const int n_bins = batch.n_bins;
    if (n_bins < 1) { set_error("bad histogram arguments"); return PISAB_ERR_ARG; }
    const bool large = n_bins > PISAB_DET_MAX_BINS;
    for (int c = 0; c < batch.n_containers; ++c) {
        const FusedContainer<IO> &C = batch.c[c];
        const bool needs_flux = !(C.flags & PISAB_CONTAINER_FLUX_SYS);
        if (C.n < 0 || (C.n > 0 && (!C.energy || !C.coszen || (needs_flux && !C.nu_flux) || !C.weights_in || !C.index))) {
            set_error("container %d: bad event arrays", c);
            return PISAB_ERR_ARG;
        }
        if (C.n > 2147483647LL) { set_error("at most 2^31-1 events per container (32-bit event indices)"); return PISAB_ERR_ARG; }
        if (!C.d_nubar && C.nubar != 1 && C.nubar != -1) { set_error("nubar must be +1 or -1"); return PISAB_ERR_ARG; }
        if (!C.d_flav && (C.flav < 0 || C.flav > 2)) { set_error("flav must be 0, 1 or 2"); return PISAB_ERR_ARG; }
    }
    const int64_t need = pisab_reweight_batch_workspace_bytes(batch.n_containers, n_bins);
    if (workspace_bytes < need || !d_workspace) {
        set_error("workspace too small: need %lld bytes", (long long)need);
        return PISAB_ERR_WORKSPACE;
    }
    OscTable ot;
    EarthTable et;
    int rc = build_osc_table(consts, &ot);
    if (rc) return rc;
    rc = build_earth_table(earth, &et);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    // standard matter potential (no NSI) -> the specialised instantiation (see H0Reg)
#ifdef PISAB_NO_STD
    const bool std_matter = false;
#else
    const bool std_matter = ot.std_matter != 0.0;
#endif
    if (consts->decay_flag == 1) {
        // Neutrino decay: reweight_hist_decay_kernel, then the common reduction launches.  One block per SM; ranks per
        // container as in the standard path's multi-wave rule, capped by the partials' workspace.
        if (large) {
            set_error("neutrino decay with more than %d bins: use propagate_earth + hist_accumulate", PISAB_DET_MAX_BINS);
            return PISAB_ERR_UNSUPPORTED;
        }
        bool any_flux = false, all_plain = true;
        for (int c = 0; c < batch.n_containers; ++c) {
            const FusedContainer<IO> &C = batch.c[c];
            all_plain = all_plain && !C.d_nubar && !C.d_flav && !C.weights_out && !C.prob_e && !C.prob_mu && !C.astro;
            if (!(C.flags & PISAB_CONTAINER_FLUX_SYS)) continue;
            any_flux = true;
            if (!C.flux_terms || !C.nu_nom || !C.nubar_nom || (uintptr_t)C.flux_terms % 32 != 0 ||
                (uintptr_t)C.nu_nom % (2 * sizeof(IO)) != 0 || (uintptr_t)C.nubar_nom % (2 * sizeof(IO)) != 0) {
                set_error("container %d: PISAB_CONTAINER_FLUX_SYS needs d_flux_terms (32-byte aligned) and both nominal fluxes", c);
                return PISAB_ERR_ARG;
            }
        }
        if (any_flux && (!all_plain || !has_sys)) {
            set_error(has_sys ? "flux.barr_simple inside the template kernel: only without per-event outputs / astro_weights"
                              : "PISAB_CONTAINER_FLUX_SYS without a pisab_flux_sys_t");
            return has_sys ? PISAB_ERR_UNSUPPORTED : PISAB_ERR_ARG;
        }
        DecayTable dt;
        rc = build_decay_table(consts, &dt);
        if (rc) return rc;
        auto kernel = reweight_hist_decay_kernel<IO>;
        const size_t smem = fused_smem_bytes<IO>(n_bins, false, false, true);
        PISAB_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kBlock, smem) != cudaSuccess || occ < 1) occ = 1;
        const int sms = (sm_count() > 0 ? sm_count() : 148) * occ;
        const int nc = batch.n_containers;
        int64_t r = (n_max + (int64_t)kBlock * 8 - 1) / ((int64_t)kBlock * 8);
        const int64_t cap = (int64_t)sms * 8 / nc > 0 ? (int64_t)sms * 8 / nc : 1;
        if (r > cap) r = cap;
        const int64_t fill = sms / nc > 0 ? sms / nc : 1, by_thread = (n_max + kBlock - 1) / kBlock;
        if (r < fill) r = by_thread < fill ? by_thread : fill;
        if (r < 1) r = 1;
        const int ranks = (int)r;
        {
            LaunchTimer t(s);
            kernel<<<ranks * nc, kBlock, smem, s>>>(ot, dt, et, batch, ranks, (double *)d_workspace, FusedEpi{});
            note_launch();
        }
        PISAB_CUDA_CHECK(cudaGetLastError());
        if (d_batch_out && epi) {
            unsigned *d_arrive = epilogue_counter(s);
            if (!d_arrive) { set_error("could not allocate the epilogue arrival counter"); return PISAB_ERR_CUDA; }
            return hist_reduce_chi2((const double *)d_workspace, ranks, n_bins, nc, epi->d_bin_scales, epi->d_observed,
                                    d_batch_out, epi->d_total, epi->d_chi2, d_arrive, s);
        }
        if (d_batch_out) return hist_reduce_batch((const double *)d_workspace, ranks, n_bins, nc, d_batch_out, s);
        return hist_reduce_partials((const double *)d_workspace, ranks, n_bins, d_hist, d_hist_w2, s);
    }
    const bool mp = sizeof(IO) == 4 && f32_math_mixed();
    // (large binnings keep no bins in shared memory, only warp_fixed_add's [4][32] words per warp: the size of a
    // 48-bin WarpHist)
    const size_t smem_events = fused_smem_bytes<IO>(large ? 48 : n_bins, std_matter, mp);
    bool plain = true;
    for (int c = 0; c < batch.n_containers; ++c) {
        const FusedContainer<IO> &C = batch.c[c];
        plain = plain && !C.d_nubar && !C.d_flav && !C.weights_out && !C.prob_e && !C.prob_mu && !C.astro;
    }
    // FP32 mode with pair-aligned containers: two events per thread (reweight_hist_pair_kernel)
    bool pairs = mp && plain && !large && sizeof(IO) == 4;
    for (int c = 0; c < batch.n_containers && pairs; ++c) {
        const FusedContainer<IO> &C = batch.c[c];
        pairs = (C.flags & PISAB_CONTAINER_PAIR_ALIGNED) && (C.n % 2 == 0) && !C.order;
    }
    bool flux = false;
    for (int c = 0; c < batch.n_containers; ++c) {
        const FusedContainer<IO> &C = batch.c[c];
        if (!(C.flags & PISAB_CONTAINER_FLUX_SYS)) continue;
        flux = true;
        if (!C.flux_terms || !C.nu_nom || !C.nubar_nom || (uintptr_t)C.flux_terms % 32 != 0 ||
            (uintptr_t)C.nu_nom % (2 * sizeof(IO)) != 0 || (uintptr_t)C.nubar_nom % (2 * sizeof(IO)) != 0) {
            set_error("container %d: PISAB_CONTAINER_FLUX_SYS needs d_flux_terms (32-byte aligned) and both nominal fluxes", c);
            return PISAB_ERR_ARG;
        }
    }
    if (flux && (!plain || large || !has_sys)) {
        set_error(has_sys ? "flux.barr_simple inside the template kernel: only without per-event outputs / astro_weights and up to %d bins"
                          : "PISAB_CONTAINER_FLUX_SYS without a pisab_flux_sys_t (up to %d bins)", PISAB_DET_MAX_BINS);
        return has_sys ? PISAB_ERR_UNSUPPORTED : PISAB_ERR_ARG;
    }
    if (large && (!plain || !d_batch_out)) {
        set_error("more than %d bins: only the batched form without per-event outputs is fused; use propagate_earth + "
                  "hist_accumulate", PISAB_DET_MAX_BINS);
        return PISAB_ERR_UNSUPPORTED;
    }
    auto kernel = large ? (std_matter ? fused_kernel_large<IO, true>(mp) : fused_kernel_large<IO, false>(mp))
                        : std_matter ? (plain ? fused_kernel<IO, true, true>(mp) : fused_kernel<IO, true, false>(mp))
                                     : (plain ? fused_kernel<IO, false, true>(mp) : fused_kernel<IO, false, false>(mp));
    if (flux) kernel = std_matter ? fused_kernel_flux<IO, true>(mp) : fused_kernel_flux<IO, false>(mp);
    if (pairs) {
        if constexpr (sizeof(IO) == 4) {
            kernel = flux ? (std_matter ? reweight_hist_pair_kernel<true, true> : reweight_hist_pair_kernel<false, true>)
                          : (std_matter ? reweight_hist_pair_kernel<true, false> : reweight_hist_pair_kernel<false, false>);
        }
    }
    size_t smem = pairs ? fused_pair_smem_bytes(n_bins) : smem_events;
    const int64_t n_units = pairs ? (n_max + 1) / 2 : n_max;   // what a thread iterates over: events or pairs
    const int64_t unit_min = pairs ? 4 : 8;                    // >= 8 events per thread
    int block = pairs ? kPairBlock : kBlock;
    // ranks (blocks) per container: at most PISAB_SPREAD_WAVES waves of resident blocks in total (blocks of
    // different containers differ in cost, so they are kept short: 1/32 of the run each; 4 waves cost 12 %, 8
    // waves 5 %, 32 and 64 are level -- profiles/r01_fused_kernel_variants.txt) with >= 8 events per thread;
    // small containers instead get up to one block per 256 events as long as one resident wave holds them all
    // (latency of one or two events per template)
    int ranks = 1, resident_blocks = 1;
    bool single_wave = false; // the whole grid is resident at once: spread the expensive events over the SMs (interleave)
    auto plan = [&]() -> int {
        {
            // static (tables) + dynamic (histogram, per-thread state and staging) exceed the 48 KB default
            cudaFuncAttributes fa;
            PISAB_CUDA_CHECK(cudaFuncGetAttributes(&fa, kernel));
            int dev = 0, optin = 0;
            PISAB_CUDA_CHECK(cudaGetDevice(&dev));
            PISAB_CUDA_CHECK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
            if (fa.sharedSizeBytes + smem > (size_t)optin) {
                set_error("fused reweight+hist: %d bins need %zu bytes of shared memory per block (limit %d); "
                          "use propagate_earth + hist_accumulate", n_bins, fa.sharedSizeBytes + smem, optin);
                return PISAB_ERR_UNSUPPORTED;
            }
            if (fa.sharedSizeBytes + smem > 48 * 1024)
                PISAB_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, block, smem) != cudaSuccess || occ < 1) occ = 1;
        const int resident = (sm_count() > 0 ? sm_count() : 148) * occ;
        const int nc = batch.n_containers;
        const int64_t by_work = (n_units + (int64_t)block * unit_min - 1) / ((int64_t)block * unit_min);
        const int64_t by_thread = (n_units + block - 1) / block;
        int64_t r = by_work;
        // (rounded down: 790 x 12 = 9480 blocks on 296 slots would leave 8 blocks for a 33rd wave)
        // (measured flat between 16 and 64 waves once the count is whole: 9.29 / 9.33 / 9.31 / 9.30 / 9.27e9 events/s at
        // 16 / 24 / 32 / 48 / 64; the rounding itself is worth 1.1 %)
        const int64_t cap = (int64_t)resident * PISAB_SPREAD_WAVES / nc > 0 ? (int64_t)resident * PISAB_SPREAD_WAVES / nc : 1;
        if (r > cap) r = cap;
        // one resident wave spread over the containers -- rounded DOWN: a grid of 300 blocks on 296 resident slots
        // runs as two waves and doubles the time of an analysis-size template
        const int64_t fill = resident / nc > 0 ? resident / nc : 1;
        if (r < fill) r = by_thread < fill ? by_thread : fill;
        int64_t ws_cap = (int64_t)(sm_count() > 0 ? sm_count() : 148) * 16; // pisab_hist_workspace_bytes
        if (r > ws_cap) r = ws_cap;
        if (r < 1) r = 1;
        ranks = (int)r;
        resident_blocks = resident;
        single_wave = (int64_t)ranks * nc <= resident;
        return PISAB_OK;
    };
    rc = plan();
    if (rc) return rc;
    if (pairs && std_matter && (int64_t)ranks * batch.n_containers >= 8 * (int64_t)resident_blocks) {
        // long multi-wave grid of the standard-matter pair kernel: 64 threads x 8 blocks per SM (see kPairBlockSmall;
        // a two-wave grid -- 1.2e6 events -- is 2 % faster with the 128-thread blocks)
        if constexpr (sizeof(IO) == 4) {
            const auto kernel0 = kernel;
            const int block0 = block, ranks0 = ranks;
            const size_t smem0 = smem;
            kernel = flux ? reweight_hist_pair_kernel<true, true, false, kPairBlockSmall>
                          : reweight_hist_pair_kernel<true, false, false, kPairBlockSmall>;
            block = kPairBlockSmall;
            smem = fused_pair_smem_bytes(n_bins, block);
            rc = plan();
            if (rc) return rc;
            if (single_wave) { // (twice the resident blocks could hold it after all: stay with the 128-thread form)
                kernel = kernel0; block = block0; smem = smem0; ranks = ranks0; single_wave = false;
            }
        }
    }
    if (pairs && single_wave) {
        // the instantiation with the interleaved chunk order and the in-kernel epilogue (same registers and shared memory)
        if constexpr (sizeof(IO) == 4) {
            kernel = flux ? (std_matter ? reweight_hist_pair_kernel<true, true, true> : reweight_hist_pair_kernel<false, true, true>)
                          : (std_matter ? reweight_hist_pair_kernel<true, false, true> : reweight_hist_pair_kernel<false, false, true>);
            PISAB_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
    }
    const int grid = ranks * batch.n_containers;
    if (large) {
        // workspace: [bounds: one u64 per container, padded to 256 B] [FixedAcc containers x 2 x n_bins]
        unsigned long long *d_bounds = (unsigned long long *)d_workspace;
        FixedAcc *d_acc = (FixedAcc *)((char *)d_workspace + 256);
        const size_t acc_bytes = sizeof(FixedAcc) * (size_t)batch.n_containers * 2 * n_bins;
        PISAB_CUDA_CHECK(cudaMemsetAsync(d_workspace, 0, 256 + acc_bytes, s));
        int64_t bpc64 = (n_max + 256 * 64 - 1) / (256 * 64); // >= 64 events per thread, at most 1024 blocks per container
        const int bpc = (int)(bpc64 < 1 ? 1 : (bpc64 > 1024 ? 1024 : bpc64));
        fused_bound_kernel<IO><<<bpc * batch.n_containers, 256, 0, s>>>(batch, bpc, d_bounds);
        note_launch();
        {
            LaunchTimer t(s);
            kernel<<<grid, kBlock, smem, s>>>(ot, et, batch, ranks, (double *)d_acc, d_bounds, single_wave ? 1 : 0, FusedEpi{});
            note_launch();
        }
        const int total = batch.n_containers * 2 * n_bins;
        fused_fixed_finish_kernel<IO><<<(total + 255) / 256, 256, 0, s>>>(batch, d_acc, d_bounds, d_batch_out);
        note_launch();
        PISAB_CUDA_CHECK(cudaGetLastError());
        return PISAB_OK;
    }
    // Single-wave grids (analysis-size samples, where a launch is a visible fraction of a template) end inside the
    // template kernel (fused_epilogue): reduction of the partial histograms, and for pisab_reweight_hist_chi2_* the
    // per-bin scales, the container sum and the chi2 -- ONE launch per hypothesis.  Multi-wave grids keep the second
    // launch: there the last block of a container would sum thousands of partial histograms alone (+0.06 ms on the
    // 10.9 ms of 1e8 events in FP64, +0.2 ms on 7.3 ms in FP32 mode; scratch/r02_run19.sh).
    FusedEpi fe = {};
    if (d_batch_out && single_wave) {
        unsigned *d_arrive = epilogue_counter(s);
        const size_t scratch = (size_t)(2 * n_bins + block) * sizeof(double);
        if (d_arrive && batch.n_containers < 63 && scratch <= smem) {
            fe.out = d_batch_out;
            fe.arrive = d_arrive;
            if (epi) { fe.bin_scales = epi->d_bin_scales; fe.observed = epi->d_observed; fe.total = epi->d_total; fe.chi2 = epi->d_chi2; }
        }
    }
    {
        LaunchTimer t(s);
        // bit 0: 32-event chunks round-robin over the blocks, bit 1: odd rounds backwards (fused_template_body)
        kernel<<<grid, block, smem, s>>>(ot, et, batch, ranks, (double *)d_workspace, nullptr, single_wave ? 3 : 0, fe);
        note_launch();
    }
    PISAB_CUDA_CHECK(cudaGetLastError());
    if (fe.out) return PISAB_OK;
    if (d_batch_out && epi) {
        unsigned *d_arrive = epilogue_counter(s);
        if (!d_arrive) { set_error("could not allocate the epilogue arrival counter"); return PISAB_ERR_CUDA; }
        return hist_reduce_chi2((const double *)d_workspace, ranks, n_bins, batch.n_containers, epi->d_bin_scales,
                                epi->d_observed, d_batch_out, epi->d_total, epi->d_chi2, d_arrive, s);
    }
    if (d_batch_out)
        return hist_reduce_batch((const double *)d_workspace, ranks, n_bins, batch.n_containers, d_batch_out, s);
    return hist_reduce_partials((const double *)d_workspace, ranks, n_bins, d_hist, d_hist_w2, s);
}

template <typename IO>
static int reweight_hist_impl(const pisab_osc_consts_t *consts, const pisab_earth_t *earth,
                              int32_t nubar, const int32_t *d_nubar, int32_t flav,
                              const int32_t *d_flav, const IO *d_energy, const IO *d_coszen,
                              const IO *d_nu_flux, const IO *d_weights_in, const int32_t *d_index,
                              const int32_t *d_order, int64_t n, int32_t n_bins, double *d_hist,
                              double *d_hist_w2,
                              IO *d_weights_out, IO *d_prob_e, IO *d_prob_mu, void *d_workspace,
                              int64_t workspace_bytes, void *stream) {
    if (!d_hist) { set_error("bad histogram arguments"); return PISAB_ERR_ARG; }
    FusedBatch<IO> batch = {};
    batch.n_containers = 1;
    batch.n_bins = n_bins;
    FusedContainer<IO> &C = batch.c[0];
    C.energy = d_energy; C.coszen = d_coszen; C.nu_flux = d_nu_flux; C.weights_in = d_weights_in;
    C.index = d_index; C.order = d_order; C.d_nubar = d_nubar; C.d_flav = d_flav;
    C.weights_out = d_weights_out; C.prob_e = d_prob_e; C.prob_mu = d_prob_mu;
    C.n = n; C.scale = 1.0; C.nubar = nubar; C.flav = flav;
    return reweight_hist_batch_impl<IO>(consts, earth, batch, n, nullptr, d_hist, d_hist_w2, d_workspace,
                                        workspace_bytes, stream);
}

template <typename IO>
static int reweight_hist_batch_abi(const pisab_osc_consts_t *consts, const pisab_earth_t *earth,
                                   const pisab_container_t *containers, int32_t n_containers,
                                   int32_t n_bins, const pisab_flux_sys_t *flux_sys, double *d_hist,
                                   void *d_workspace, int64_t workspace_bytes, void *stream,
                                   const Chi2Epilogue *epi = nullptr) {
    if (!containers || n_containers < 1 || n_containers > PISAB_MAX_BATCH || !d_hist) {
        set_error("n_containers must be in [1, %d] and the output non-null", PISAB_MAX_BATCH);
        return PISAB_ERR_ARG;
    }
    FusedBatch<IO> batch = {};
    batch.n_containers = n_containers;
    batch.n_bins = n_bins;
    if (flux_sys)
        batch.sys = BarrSys{flux_sys->nue_numu_ratio, flux_sys->nu_nubar_ratio, flux_sys->delta_index,
                            flux_sys->barr_uphor_ratio, flux_sys->barr_nu_nubar_ratio};
    int64_t n_max = 0;
    for (int c = 0; c < n_containers; ++c) {
        const pisab_container_t &S = containers[c];
        FusedContainer<IO> &C = batch.c[c];
        C.energy = (const IO *)S.d_energy; C.coszen = (const IO *)S.d_coszen;
        C.nu_flux = (const IO *)S.d_nu_flux; C.weights_in = (const IO *)S.d_weights;
        C.index = S.d_index; C.order = S.d_order; C.d_nubar = nullptr; C.d_flav = nullptr;
        C.weights_out = (IO *)S.d_weights_out; C.prob_e = nullptr; C.prob_mu = nullptr;
        C.n = S.n; C.scale = S.scale; C.nubar = S.nubar; C.flav = S.flav; C.flags = S.flags;
        C.flux_terms = S.d_flux_terms; C.nu_nom = (const IO *)S.d_nu_flux_nominal;
        C.nubar_nom = (const IO *)S.d_nubar_flux_nominal;
        C.astro = (const IO *)S.d_astro_weights;
        if (S.n > n_max) n_max = S.n;
    }
    if (epi && n_bins > PISAB_DET_MAX_BINS) {
        set_error("the chi2 epilogue supports up to %d bins", PISAB_DET_MAX_BINS);
        return PISAB_ERR_UNSUPPORTED;
    }
    return reweight_hist_batch_impl<IO>(consts, earth, batch, n_max, d_hist, nullptr, nullptr, d_workspace,
                                        workspace_bytes, stream, epi, flux_sys != nullptr);
}

// ranks (blocks) per (template, container) of a scan: every thread should see >= 8 events of the largest
// container, and the launch should have at least one full wave of blocks
static int scan_ranks(int64_t n_max, int n_templates, int n_containers) {
    const int sms = sm_count() > 0 ? sm_count() : 148;
    const int64_t pairs = (int64_t)n_templates * n_containers;
    int64_t r = (n_max + (int64_t)kBlock * 8 - 1) / ((int64_t)kBlock * 8);
    const int64_t by_thread = (n_max + kBlock - 1) / kBlock;
    const int64_t fill = ((int64_t)sms * 2 + pairs - 1) / pairs;
    if (r < fill) r = by_thread < fill ? by_thread : fill;
    if (r < 1) r = 1;
    if (r > (int64_t)sms * 2) r = (int64_t)sms * 2;
    return (int)r;
}

template <typename IO>
static int reweight_hist_scan_abi(const pisab_osc_consts_t *consts, int32_t n_templates,
                                  const pisab_earth_t *earth, const pisab_container_t *containers,
                                  int32_t n_containers, int32_t n_bins, double *d_hist, void *d_workspace,
                                  int64_t workspace_bytes, void *stream) {
    if (!consts || n_templates < 1 || !containers || n_containers < 1 || n_containers > PISAB_MAX_BATCH || !d_hist) {
        set_error("scan needs >= 1 template, 1..%d containers and a non-null output", PISAB_MAX_BATCH);
        return PISAB_ERR_ARG;
    }
    if (n_bins < 1 || n_bins > PISAB_DET_MAX_BINS) { set_error("n_bins outside [1, %d]", PISAB_DET_MAX_BINS); return PISAB_ERR_UNSUPPORTED; }
    FusedBatch<IO> batch = {};
    batch.n_containers = n_containers;
    batch.n_bins = n_bins;
    int64_t n_max = 0;
    for (int c = 0; c < n_containers; ++c) {
        const pisab_container_t &S = containers[c];
        FusedContainer<IO> &C = batch.c[c];
        if (S.flags & PISAB_CONTAINER_FLUX_SYS) { set_error("scan: PISAB_CONTAINER_FLUX_SYS is not supported; apply the flux systematics first"); return PISAB_ERR_UNSUPPORTED; }
        if (S.d_astro_weights) { set_error("scan: astro_weights are not supported"); return PISAB_ERR_UNSUPPORTED; }
        if (S.n < 0 || S.n > 2147483647LL || (S.n > 0 && (!S.d_energy || !S.d_coszen || !S.d_nu_flux || !S.d_weights || !S.d_index))) {
            set_error("container %d: bad event arrays", c);
            return PISAB_ERR_ARG;
        }
        if ((S.nubar != 1 && S.nubar != -1) || S.flav < 0 || S.flav > 2) { set_error("container %d: bad nubar / flav", c); return PISAB_ERR_ARG; }
        C.energy = (const IO *)S.d_energy; C.coszen = (const IO *)S.d_coszen;
        C.nu_flux = (const IO *)S.d_nu_flux; C.weights_in = (const IO *)S.d_weights;
        C.index = S.d_index; C.order = S.d_order;
        C.n = S.n; C.scale = S.scale; C.nubar = S.nubar; C.flav = S.flav; // per-event outputs stay NULL
        if (S.n > n_max) n_max = S.n;
    }
    std::vector<OscTable> tables((size_t)n_templates);
    bool all_std = true;
    bool any_decay = false;
    for (int t = 0; t < n_templates; ++t) any_decay = any_decay || consts[t].decay_flag == 1;
    std::vector<DecayTable> dtables(any_decay ? (size_t)n_templates : 0);
    for (int t = 0; t < n_templates; ++t) {
        const int rc = build_osc_table(consts + t, &tables[t]);
        if (rc) return rc;
        if (any_decay) {
            // one kernel for the whole scan: templates without decay get a zero table (and no vacuum shortcut)
            tables[t].vac_ok = 0.0;
            memset(&dtables[t], 0, sizeof(DecayTable));
            if (consts[t].decay_flag == 1) {
                const int rd = build_decay_table(consts + t, &dtables[t]);
                if (rd) return rd;
            }
        }
        all_std = all_std && tables[t].std_matter != 0.0;
    }
    EarthTable et;
    const int rc = build_earth_table(earth, &et);
    if (rc) return rc;
    const int ranks = scan_ranks(n_max, n_templates, n_containers);
    const int64_t need = pisab_reweight_scan_workspace_bytes(n_templates, n_containers, n_bins, n_max);
    if (!d_workspace || workspace_bytes < need) { set_error("workspace too small: need %lld bytes", (long long)need); return PISAB_ERR_WORKSPACE; }
    cudaStream_t s = (cudaStream_t)stream;
    const size_t osc_bytes = ((size_t)n_templates * sizeof(OscTable) + 255) / 256 * 256;
    const size_t table_bytes = ((size_t)n_templates * (sizeof(OscTable) + sizeof(DecayTable)) + 511) / 256 * 256;
    OscTable *d_tables = (OscTable *)d_workspace;
    DecayTable *d_dtables = (DecayTable *)((char *)d_workspace + osc_bytes);
    double *d_partials = (double *)((char *)d_workspace + table_bytes);
    PISAB_CUDA_CHECK(cudaMemcpyAsync(d_tables, tables.data(), (size_t)n_templates * sizeof(OscTable), cudaMemcpyHostToDevice, s));
    if (any_decay) {
        PISAB_CUDA_CHECK(cudaMemcpyAsync(d_dtables, dtables.data(), (size_t)n_templates * sizeof(DecayTable), cudaMemcpyHostToDevice, s));
        auto kernel = reweight_hist_scan_decay_kernel<IO>;
        const size_t smem = fused_smem_bytes<IO>(n_bins, false, false, true);
        cudaFuncAttributes fa;
        PISAB_CUDA_CHECK(cudaFuncGetAttributes(&fa, kernel));
        int dev = 0, optin = 0;
        PISAB_CUDA_CHECK(cudaGetDevice(&dev));
        PISAB_CUDA_CHECK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        if (fa.sharedSizeBytes + smem > (size_t)optin) { set_error("scan: %d bins exceed the shared-memory budget", n_bins); return PISAB_ERR_UNSUPPORTED; }
        PISAB_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        {
            LaunchTimer t(s);
            kernel<<<n_templates * n_containers * ranks, kBlock, smem, s>>>(d_tables, d_dtables, et, batch, ranks, d_partials);
            note_launch();
        }
        PISAB_CUDA_CHECK(cudaGetLastError());
        return hist_reduce_batch(d_partials, ranks, n_bins, n_templates * n_containers, d_hist, s);
    }
    const bool mp = sizeof(IO) == 4 && f32_math_mixed();
    const size_t smem = fused_smem_bytes<IO>(n_bins, all_std, mp);
    auto kernel = all_std ? scan_kernel<IO, true>(mp) : scan_kernel<IO, false>(mp);
    {
        cudaFuncAttributes fa;
        PISAB_CUDA_CHECK(cudaFuncGetAttributes(&fa, kernel));
        int dev = 0, optin = 0;
        PISAB_CUDA_CHECK(cudaGetDevice(&dev));
        PISAB_CUDA_CHECK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        if (fa.sharedSizeBytes + smem > (size_t)optin) { set_error("scan: %d bins exceed the shared-memory budget", n_bins); return PISAB_ERR_UNSUPPORTED; }
        PISAB_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    {
        LaunchTimer t(s);
        kernel<<<n_templates * n_containers * ranks, kBlock, smem, s>>>(d_tables, et, batch, ranks, d_partials);
        note_launch();
    }
    PISAB_CUDA_CHECK(cudaGetLastError());
    return hist_reduce_batch(d_partials, ranks, n_bins, n_templates * n_containers, d_hist, s);
}

extern "C" {

int pisab_prob3_propagate_earth_f64(const pisab_osc_consts_t *consts, const pisab_earth_t *earth,
                                    int32_t nubar, const int32_t *d_nubar, int32_t flav,
                                    const int32_t *d_flav, const double *d_energy,
                                    const double *d_coszen, const int32_t *d_order, int64_t n,
                                    double *d_probability, double *d_prob_e, double *d_prob_mu,
                                    void *stream) {
    return propagate_earth_impl<double>(consts, earth, nubar, d_nubar, flav, d_flav, d_energy, d_coszen, d_order, n,
                                        d_probability, d_prob_e, d_prob_mu, stream);
}
int pisab_prob3_propagate_earth_f32(const pisab_osc_consts_t *consts, const pisab_earth_t *earth,
                                    int32_t nubar, const int32_t *d_nubar, int32_t flav,
                                    const int32_t *d_flav, const float *d_energy, const float *d_coszen,
                                    const int32_t *d_order, int64_t n, float *d_probability,
                                    float *d_prob_e, float *d_prob_mu, void *stream) {
    return propagate_earth_impl<float>(consts, earth, nubar, d_nubar, flav, d_flav, d_energy, d_coszen, d_order, n,
                                       d_probability, d_prob_e, d_prob_mu, stream);
}
int pisab_prob3_propagate_layers_f64(const pisab_osc_consts_t *consts, int32_t nubar,
                                     const int32_t *d_nubar, const double *d_energy,
                                     const double *d_densities, const double *d_distances, int64_t n,
                                     int32_t n_layers, double *d_probability, void *stream) {
    return propagate_layers_impl<double>(consts, nubar, d_nubar, d_energy, d_densities, d_distances, n,
                                         n_layers, d_probability, stream);
}
int pisab_prob3_propagate_layers_f32(const pisab_osc_consts_t *consts, int32_t nubar,
                                     const int32_t *d_nubar, const float *d_energy,
                                     const float *d_densities, const float *d_distances, int64_t n,
                                     int32_t n_layers, float *d_probability, void *stream) {
    return propagate_layers_impl<float>(consts, nubar, d_nubar, d_energy, d_densities, d_distances, n,
                                        n_layers, d_probability, stream);
}
int pisab_reweight_hist_f64(const pisab_osc_consts_t *consts, const pisab_earth_t *earth,
                            int32_t nubar, const int32_t *d_nubar, int32_t flav, const int32_t *d_flav,
                            const double *d_energy, const double *d_coszen, const double *d_nu_flux,
                            const double *d_weights_in, const int32_t *d_index,
                            const int32_t *d_order, int64_t n,
                            int32_t n_bins, double *d_hist, double *d_hist_w2, double *d_weights_out,
                            double *d_prob_e, double *d_prob_mu, void *d_workspace,
                            int64_t workspace_bytes, void *stream) {
    return reweight_hist_impl<double>(consts, earth, nubar, d_nubar, flav, d_flav, d_energy, d_coszen,
                                      d_nu_flux, d_weights_in, d_index, d_order, n, n_bins, d_hist, d_hist_w2,
                                      d_weights_out, d_prob_e, d_prob_mu, d_workspace, workspace_bytes, stream);
}
int pisab_reweight_hist_f32(const pisab_osc_consts_t *consts, const pisab_earth_t *earth,
                            int32_t nubar, const int32_t *d_nubar, int32_t flav, const int32_t *d_flav,
                            const float *d_energy, const float *d_coszen, const float *d_nu_flux,
                            const float *d_weights_in, const int32_t *d_index, const int32_t *d_order,
                            int64_t n, int32_t n_bins,
                            double *d_hist, double *d_hist_w2, float *d_weights_out, float *d_prob_e,
                            float *d_prob_mu, void *d_workspace, int64_t workspace_bytes, void *stream) {
    return reweight_hist_impl<float>(consts, earth, nubar, d_nubar, flav, d_flav, d_energy, d_coszen,
                                     d_nu_flux, d_weights_in, d_index, d_order, n, n_bins, d_hist, d_hist_w2,
                                     d_weights_out, d_prob_e, d_prob_mu, d_workspace, workspace_bytes, stream);
}

int pisab_reweight_hist_batch_f64(const pisab_osc_consts_t *consts, const pisab_earth_t *earth,
                                  const pisab_container_t *containers, int32_t n_containers,
                                  int32_t n_bins, const pisab_flux_sys_t *flux_sys, double *d_hist,
                                  void *d_workspace, int64_t workspace_bytes, void *stream) {
    return reweight_hist_batch_abi<double>(consts, earth, containers, n_containers, n_bins, flux_sys, d_hist, d_workspace,
                                           workspace_bytes, stream);
}
int pisab_reweight_hist_batch_f32(const pisab_osc_consts_t *consts, const pisab_earth_t *earth,
                                  const pisab_container_t *containers, int32_t n_containers,
                                  int32_t n_bins, const pisab_flux_sys_t *flux_sys, double *d_hist,
                                  void *d_workspace, int64_t workspace_bytes, void *stream) {
    return reweight_hist_batch_abi<float>(consts, earth, containers, n_containers, n_bins, flux_sys, d_hist, d_workspace,
                                          workspace_bytes, stream);
}

int pisab_reweight_hist_chi2_f64(const pisab_osc_consts_t *consts, const pisab_earth_t *earth,
                                 const pisab_container_t *containers, int32_t n_containers, int32_t n_bins,
                                 const pisab_flux_sys_t *flux_sys, const double *d_bin_scales, const double *d_observed, double *d_hist, double *d_total,
                                 double *d_chi2, void *d_workspace, int64_t workspace_bytes, void *stream) {
    const Chi2Epilogue epi{d_bin_scales, d_observed, d_total, d_chi2};
    return reweight_hist_batch_abi<double>(consts, earth, containers, n_containers, n_bins, flux_sys, d_hist, d_workspace,
                                           workspace_bytes, stream, &epi);
}
int pisab_reweight_hist_chi2_f32(const pisab_osc_consts_t *consts, const pisab_earth_t *earth,
                                 const pisab_container_t *containers, int32_t n_containers, int32_t n_bins,
                                 const pisab_flux_sys_t *flux_sys, const double *d_bin_scales, const double *d_observed, double *d_hist, double *d_total,
                                 double *d_chi2, void *d_workspace, int64_t workspace_bytes, void *stream) {
    const Chi2Epilogue epi{d_bin_scales, d_observed, d_total, d_chi2};
    return reweight_hist_batch_abi<float>(consts, earth, containers, n_containers, n_bins, flux_sys, d_hist, d_workspace,
                                          workspace_bytes, stream, &epi);
}

int pisab_hist_scale_sum_chi2(const double *d_partials, int32_t n_blocks, int32_t n_containers, int32_t n_bins,
                              const double *d_bin_scales, const double *d_observed, double *d_hist, double *d_total,
                              double *d_chi2, void *stream) {
    if (!d_partials || !d_hist || n_blocks < 1 || n_containers < 1 || n_containers > PISAB_MAX_BATCH || n_bins < 1 ||
        n_bins > PISAB_DET_MAX_BINS || (d_observed && !d_chi2)) {
        set_error("hist_scale_sum_chi2: bad arguments (1..%d containers, 1..%d bins)", PISAB_MAX_BATCH, PISAB_DET_MAX_BINS);
        return PISAB_ERR_ARG;
    }
    unsigned *d_arrive = epilogue_counter((cudaStream_t)stream);
    if (!d_arrive) { set_error("could not allocate the epilogue arrival counter"); return PISAB_ERR_CUDA; }
    return hist_reduce_chi2(d_partials, n_blocks, n_bins, n_containers, d_bin_scales, d_observed, d_hist, d_total, d_chi2,
                            d_arrive, (cudaStream_t)stream);
}

int64_t pisab_reweight_scan_workspace_bytes(int32_t n_templates, int32_t n_containers, int32_t n_bins, int64_t n_max) {
    if (n_templates < 1) n_templates = 1;
    if (n_containers < 1) n_containers = 1;
    const int ranks = scan_ranks(n_max, n_templates, n_containers);
    // (per template: OscTable + DecayTable, each block rounded to 256 bytes -- see reweight_hist_scan_abi)
    const int64_t table_bytes = ((int64_t)n_templates * (int64_t)(sizeof(OscTable) + sizeof(DecayTable)) + 511) / 256 * 256;
    return table_bytes + (int64_t)n_templates * n_containers * ranks * 2 * (int64_t)n_bins * (int64_t)sizeof(double);
}
int pisab_reweight_hist_scan_f64(const pisab_osc_consts_t *consts, int32_t n_templates, const pisab_earth_t *earth,
                                 const pisab_container_t *containers, int32_t n_containers, int32_t n_bins,
                                 double *d_hist, void *d_workspace, int64_t workspace_bytes, void *stream) {
    return reweight_hist_scan_abi<double>(consts, n_templates, earth, containers, n_containers, n_bins, d_hist,
                                          d_workspace, workspace_bytes, stream);
}
int pisab_reweight_hist_scan_f32(const pisab_osc_consts_t *consts, int32_t n_templates, const pisab_earth_t *earth,
                                 const pisab_container_t *containers, int32_t n_containers, int32_t n_bins,
                                 double *d_hist, void *d_workspace, int64_t workspace_bytes, void *stream) {
    return reweight_hist_scan_abi<float>(consts, n_templates, earth, containers, n_containers, n_bins, d_hist,
                                         d_workspace, workspace_bytes, stream);
}

} // extern "C"
