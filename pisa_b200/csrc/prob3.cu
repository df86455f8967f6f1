// prob3.cu -- propagation kernels + their C-ABI entry points (include/pisa_b200.h).
#include <math.h>
#include <string.h>

#include "hist_device.cuh"
#include "prob3_device.cuh"

namespace pisab {

// ---------------------------------------------------------------------------------------------
// host: parameter tables
// ---------------------------------------------------------------------------------------------
static void pack_herm(const double m[3][3][2], double scale, Herm3 *h) {
    h->d0 = scale * m[0][0][0];
    h->d1 = scale * m[1][1][0];
    h->d2 = scale * m[2][2][0];
    h->r01 = scale * m[0][1][0];
    h->i01 = scale * m[0][1][1];
    h->r02 = scale * m[0][2][0];
    h->i02 = scale * m[0][2][1];
    h->r12 = scale * m[1][2][0];
    h->i12 = scale * m[1][2][1];
}

static bool is_hermitian(const double m[3][3][2]) {
    double scale = 0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) scale = fmax(scale, fmax(fabs(m[i][j][0]), fabs(m[i][j][1])));
    const double tol = 1e-12 * scale + 1e-300;
    for (int i = 0; i < 3; ++i) {
        if (fabs(m[i][i][1]) > tol) return false;
        for (int j = i + 1; j < 3; ++j)
            if (fabs(m[i][j][0] - m[j][i][0]) > tol || fabs(m[i][j][1] + m[j][i][1]) > tol) return false;
    }
    return true;
}

int build_osc_table(const pisab_osc_consts_t *c, OscTable *out) {
    if (!c || !out) { set_error("null osc consts"); return PISAB_ERR_ARG; }
    if (c->decay_flag == 1) {
        // numba_osc_kernels.py:445-451 -> get_dms_numerical (numpy.linalg.eigvals): out of scope
        set_error("decay_flag == 1 (neutrino decay) is not supported by the B200 path");
        return PISAB_ERR_UNSUPPORTED;
    }
    double U[3][3][2], V[3][3][2], Lr[3][3][2], Hv[3][3][2];
    memcpy(U, c->mix, sizeof U);
    memcpy(V, c->mat_pot, sizeof V);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { Lr[i][j][0] = c->lri_pot[i * 3 + j]; Lr[i][j][1] = 0.0; }
    if (!is_hermitian(V)) { set_error("mat_pot must be Hermitian"); return PISAB_ERR_UNSUPPORTED; }
    if (!is_hermitian(Lr)) { set_error("lri_pot must be symmetric"); return PISAB_ERR_UNSUPPORTED; }
    // H_vac = U diag(0, dm[1][0], dm[2][0]) U^dagger  (get_H_vac, numba_osc_kernels.py:534-569)
    const double d[3] = {0.0, c->dm[1 * 3 + 0], c->dm[2 * 3 + 0]};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double re = 0, im = 0;
            for (int k = 0; k < 3; ++k) {
                // U[i][k] * d[k] * conj(U[j][k])
                const double ar = U[i][k][0], ai = U[i][k][1], br = U[j][k][0], bi = -U[j][k][1];
                re += d[k] * (ar * br - ai * bi);
                im += d[k] * (ar * bi + ai * br);
            }
            Hv[i][j][0] = re;
            Hv[i][j][1] = im;
        }
    pack_herm(Hv, 0.5, &out->hv[0]);   // one_over_two_e = 0.5 / energy (:443)
    pack_herm(Hv, -0.5, &out->hv[1]);  // antineutrinos, see common.cuh
    pack_herm(V, 0.5 * 1.52588e-4, &out->vm); // a = 0.5 * rho * tworttwoGf (:636-637)
    pack_herm(Lr, 1e9, &out->lr);      // eV -> eV^2/GeV (:438)
    return PISAB_OK;
}

int build_earth_table(const pisab_earth_t *e, EarthTable *out) {
    if (!e || !out) { set_error("null earth"); return PISAB_ERR_ARG; }
    if (e->n_radii < 2 || e->n_radii > PISAB_MAX_RADII) {
        set_error("n_radii = %d outside [2, %d]", e->n_radii, PISAB_MAX_RADII);
        return PISAB_ERR_ARG;
    }
    memset(out, 0, sizeof *out);
    out->n_radii = e->n_radii;
    out->r_det = e->r_detector;
    out->rd2 = e->r_detector * e->r_detector;
    int idx = -1;
    for (int j = 0; j < e->n_radii; ++j) {
        out->rj2[j] = e->radii[j] * e->radii[j];
        out->rho[j] = e->rho_e[j];
        out->limit[j] = e->coszen_limit[j];
        if (idx < 0 && e->radii[j] < e->r_detector) idx = j;
    }
    if (idx < 1) { set_error("no Earth shell below the detector"); return PISAB_ERR_UNSUPPORTED; }
    if (idx != 2) {
        // extCalcLayers pairs 2K - idx segments with 2K - 2 densities (layers.py:128-158); for
        // idx != 2 the reference reads out of bounds for every up-going direction.
        set_error("detector must sit inside the outermost Earth shell (first inner shell index %d != 2); "
                  "the reference's extCalcLayers is undefined for this geometry", idx);
        return PISAB_ERR_UNSUPPORTED;
    }
    out->idx_first_inner = idx;
    return PISAB_OK;
}

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

// One thread per event, layers computed in-kernel from coszen.
//   FULL: probability[n,3,3] ; otherwise prob_e / prob_mu of the event's final flavour.
template <typename IO, bool FULL>
__global__ void __launch_bounds__(kBlock, PISAB_MIN_BLOCKS)
prob3_earth_kernel(const __grid_constant__ OscTable osc, const __grid_constant__ EarthTable earth,
                   int nubar, const int32_t *__restrict__ d_nubar, int flav,
                   const int32_t *__restrict__ d_flav, const IO *__restrict__ energy,
                   const IO *__restrict__ coszen, const int32_t *__restrict__ order, int64_t n,
                   IO *__restrict__ probability, IO *__restrict__ prob_e, IO *__restrict__ prob_mu) {
    __shared__ OscTable s_osc;
    __shared__ EarthTable s_earth;
    copy_tables(osc, earth, &s_osc, &s_earth);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
        // `order` (optional) lists the events grouped by number of crossed shells so that the 32
        // lanes of a warp walk the same number of layers; results go back to the event's own slot
        const int64_t i = order ? (int64_t)__ldg(order + t) : t;
        const double e = ld(energy, i), cz = ld(coszen, i);
        const int nb = d_nubar ? __ldg(d_nubar + i) : nubar;
        const int fl = d_flav ? __ldg(d_flav + i) : flav;
        const Herm3 h0 = herm_axpy(rcp_fast(e), s_osc.hv[nb > 0 ? 0 : 1], s_osc.lr);
        if (FULL) {
            Propagator<3, 3> P;
            propagate_earth<3, 3>(h0, s_osc.vm, s_earth, cz, 0, P);
            IO *o = probability + i * 9;
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) o[a * 3 + b] = (IO)P.prob(b, a); // P(a->b) = |A[b][a]|^2
        } else {
            Propagator<1, 2> P;
            propagate_earth<1, 2>(h0, s_osc.vm, s_earth, cz, fl, P);
            prob_e[i] = (IO)P.prob(0, 0);
            prob_mu[i] = (IO)P.prob(0, 1);
        }
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
        const Herm3 h0 = herm_axpy(rcp_fast(e), s_osc.hv[nb > 0 ? 0 : 1], s_osc.lr);
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

// Fused template evaluation: probabilities + reweighting + weighted histogram (w, w^2).
template <typename IO>
__global__ void __launch_bounds__(kBlock, PISAB_MIN_BLOCKS)
reweight_hist_kernel(const __grid_constant__ OscTable osc, const __grid_constant__ EarthTable earth,
                     int nubar, const int32_t *__restrict__ d_nubar, int flav,
                     const int32_t *__restrict__ d_flav, const IO *__restrict__ energy,
                     const IO *__restrict__ coszen, const IO *__restrict__ nu_flux,
                     const IO *__restrict__ weights_in, const int32_t *__restrict__ index,
                     const int32_t *__restrict__ order, int64_t n,
                     int n_bins, double *__restrict__ partials, IO *__restrict__ weights_out,
                     IO *__restrict__ prob_e, IO *__restrict__ prob_mu) {
    extern __shared__ double s_hist[]; // [warps][2][n_bins] private bins, then staging
    __shared__ OscTable s_osc;
    __shared__ EarthTable s_earth;
    WarpHist wh(s_hist, n_bins);
    wh.clear();
    copy_tables(osc, earth, &s_osc, &s_earth);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // warp-uniform trip count so that the warp-collective histogram step is always converged
    const int64_t warp_first = first - (threadIdx.x & 31);
    for (int64_t base = warp_first; base < n; base += stride) {
        const int64_t t = base + (threadIdx.x & 31);
        double w = 0.0;
        int bin = -1;
        if (t < n) {
            const int64_t i = order ? (int64_t)__ldg(order + t) : t;
            const double e = ld(energy, i), cz = ld(coszen, i);
            const int nb = d_nubar ? __ldg(d_nubar + i) : nubar;
            const int fl = d_flav ? __ldg(d_flav + i) : flav;
            const Herm3 h0 = herm_axpy(rcp_fast(e), s_osc.hv[nb > 0 ? 0 : 1], s_osc.lr);
            Propagator<1, 2> P;
            propagate_earth<1, 2>(h0, s_osc.vm, s_earth, cz, fl, P);
            const double pe = P.prob(0, 0), pmu = P.prob(0, 1);
            // prob3.py:622: weights *= (flux_e * prob_e) + (flux_mu * prob_mu)
            const double fe = ld(nu_flux, 2 * i), fm = ld(nu_flux, 2 * i + 1);
            w = ld(weights_in, i) * (fe * pe + fm * pmu);
            bin = __ldg(index + i);
            if (weights_out) weights_out[i] = (IO)w;
            if (prob_e) prob_e[i] = (IO)pe;
            if (prob_mu) prob_mu[i] = (IO)pmu;
        }
        wh.add(bin, w);
    }
    wh.flush(partials + (size_t)blockIdx.x * 2 * n_bins);
}

} // namespace pisab

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
using namespace pisab;

static int grid_for(int64_t n, int blocks_per_sm) {
    const int sms = sm_count();
    int64_t want = (n + kBlock - 1) / kBlock;
    int64_t cap = (int64_t)(sms > 0 ? sms : 148) * blocks_per_sm;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

// Persistent grid: exactly the number of blocks that are resident at once (one wave), so that the
// grid-stride loop gives every block the same share of every layer-count class.
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
    if (d_probability) {
        LaunchTimer t(s);
        prob3_earth_kernel<IO, true><<<resident_grid(prob3_earth_kernel<IO, true>, n, 0), kBlock, 0, s>>>(
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
        LaunchTimer t(s);
        prob3_earth_kernel<IO, false><<<resident_grid(prob3_earth_kernel<IO, false>, n, 0), kBlock, 0, s>>>(
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
    if (n_layers < 1 || n_layers > PISAB_MAX_LAYERS) {
        set_error("n_layers = %d outside [1, %d] (numba_osc_kernels.py:227)", n_layers, PISAB_MAX_LAYERS);
        return PISAB_ERR_ARG;
    }
    if (!d_nubar && nubar != 1 && nubar != -1) { set_error("nubar must be +1 or -1"); return PISAB_ERR_ARG; }
    OscTable ot;
    int rc = build_osc_table(consts, &ot);
    if (rc) return rc;
    if (n == 0) return PISAB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    {
        LaunchTimer t(s);
        prob3_layers_kernel<IO><<<grid_for(n, 4), kBlock, 0, s>>>(ot, nubar, d_nubar, d_energy, d_densities,
                                                                  d_distances, n, n_layers, d_probability);
        note_launch();
    }
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
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
    if (n < 0 || n_bins < 1 || !d_hist) { set_error("bad histogram arguments"); return PISAB_ERR_ARG; }
    if (n > 0 && (!d_energy || !d_coszen || !d_nu_flux || !d_weights_in || !d_index)) {
        set_error("bad event arrays");
        return PISAB_ERR_ARG;
    }
    if (n_bins > PISAB_DET_MAX_BINS) {
        set_error("fused reweight+hist supports up to %d bins; use propagate_earth + hist_accumulate", PISAB_DET_MAX_BINS);
        return PISAB_ERR_UNSUPPORTED;
    }
    if (!d_nubar && nubar != 1 && nubar != -1) { set_error("nubar must be +1 or -1"); return PISAB_ERR_ARG; }
    if (!d_flav && (flav < 0 || flav > 2)) { set_error("flav must be 0, 1 or 2"); return PISAB_ERR_ARG; }
    if (workspace_bytes < pisab_hist_workspace_bytes(n, n_bins) || !d_workspace) {
        set_error("workspace too small: need %lld bytes", (long long)pisab_hist_workspace_bytes(n, n_bins));
        return PISAB_ERR_WORKSPACE;
    }
    OscTable ot;
    EarthTable et;
    int rc = build_osc_table(consts, &ot);
    if (rc) return rc;
    rc = build_earth_table(earth, &et);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = WarpHist::smem_bytes(kBlock, n_bins);
    if (smem > 48 * 1024)
        PISAB_CUDA_CHECK(cudaFuncSetAttribute(reweight_hist_kernel<IO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = resident_grid(reweight_hist_kernel<IO>, n, smem);
    {
        LaunchTimer t(s);
        reweight_hist_kernel<IO><<<grid, kBlock, smem, s>>>(ot, et, nubar, d_nubar, flav, d_flav, d_energy,
                                                            d_coszen, d_nu_flux, d_weights_in, d_index, d_order, n, n_bins,
                                                            (double *)d_workspace, d_weights_out, d_prob_e, d_prob_mu);
        note_launch();
    }
    PISAB_CUDA_CHECK(cudaGetLastError());
    return hist_reduce_partials((const double *)d_workspace, grid, n_bins, d_hist, d_hist_w2, s);
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

} // extern "C"
