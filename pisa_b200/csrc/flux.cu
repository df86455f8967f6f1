// flux.cu -- flux.barr_simple: Barr-style flux systematics per event.
//
// Reference: pisa/stages/flux/barr_simple.py:107-197 (apply_ratio_scale, spectral_index_scale,
// apply_sys_kernel) and pisa/utils/barr_parameterization.py:17-113 (LogLogParam, norm_fcn, ModFlux,
// modRatioUpHor, modRatioNuBar).  Per event:
//     nue/numu ratio (sum preserving) -> spectral index (E/E0)^delta -> nu/nubar ratio (sum preserving)
//     -> pick nu or nubar -> Barr nu/nubar modification -> Barr up/horizontal modification.
// Every LogLogParam of the reference is s * 10^(slope * log10(E) + intercept) [* exp(-E/cutoff)] with
// constants that do not depend on the event; they are folded on the host into BarrTable, so an event
// costs one log10, five 10^x, two exp(-E/cutoff), two Gaussians in coszen and one pow.  With ~12 FP64
// transcendentals per 64 bytes this kernel is FP64-bound, not HBM-bound (measured in profiles/).
#include <math.h>

#include "common.cuh"
#include "flux_device.cuh"

namespace pisab {

struct LogLog {
    double sign, slope, icpt; // sign * 10^(slope * log10(E) + icpt)
};
struct BarrTable {
    LogLog ave_mu, shape_mu, ave_e, shape_e, uphor_e;
    double nue_numu_ratio, nu_nubar_ratio, delta_index, uphor, nubar_sys;
};

static double sgn_h(double v) { return v == 0 ? 0.0 : (v >= 0 ? 1.0 : -1.0); }
// barr_parameterization.py:26-35 with the event-independent part evaluated once
static LogLog fold_loglog(double y1, double y2, double x1, double x2) {
    LogLog r;
    r.sign = sgn_h(y2);
    const double Y1 = sgn_h(y1) * log10(fabs(y1) + 0.0001);
    const double Y2 = log10(fabs(y2 + 0.0001));
    r.slope = (Y2 - Y1) / (x2 - x1);
    r.icpt = Y1 - 2. - r.slope * x1;
    return r;
}

__device__ __forceinline__ double loglog(const LogLog &p, double log10e) {
    return p.sign * exp10(fma(p.slope, log10e, p.icpt));
}
// barr_parameterization.py:37-40: A / sqrt(2 pi sigma^2) * exp(-x^2 / (2 sigma^2))
__device__ __forceinline__ double norm_fcn(double x, double A, double inv_norm, double inv_2s2) {
    return A * inv_norm * exp(-x * x * inv_2s2);
}
// barr_simple.py:107-136 with sum_constant = True
__device__ __forceinline__ void ratio_scale(double scale, double in1, double in2, double &o0, double &o1) {
    if (in1 == 0. && in2 == 0.) { o0 = 0.; o1 = 0.; return; }
    const double orig_ratio = in1 / in2;
    const double nw = (in1 + in2) / (1. + scale * orig_ratio);
    o0 = scale * orig_ratio * nw;
    o1 = nw;
}

template <typename IO>
__global__ void __launch_bounds__(256)
flux_barr_simple_kernel(const __grid_constant__ BarrTable T, const IO *__restrict__ energy,
                        const IO *__restrict__ coszen, const IO *__restrict__ nu_nom,
                        const IO *__restrict__ nubar_nom, int nubar, int64_t n, IO *__restrict__ out) {
    const double inv_norm36 = 1.0 / sqrt(2 * M_PI * 0.36 * 0.36), inv_2s36 = 1.0 / (2 * 0.36 * 0.36);
    const double inv_norm35 = 1.0 / sqrt(2 * M_PI * 0.35 * 0.35), inv_2s35 = 1.0 / (2 * 0.35 * 0.35);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double e = (double)__ldg(energy + i), cz = (double)__ldg(coszen + i);
        double nu0 = (double)__ldg(nu_nom + 2 * i), nu1 = (double)__ldg(nu_nom + 2 * i + 1);
        double nb0 = (double)__ldg(nubar_nom + 2 * i), nb1 = (double)__ldg(nubar_nom + 2 * i + 1);
        // nue/numu ratio, spectral index (apply_sys_kernel :158-176)
        double a0, a1, b0, b1;
        ratio_scale(T.nue_numu_ratio, nu0, nu1, a0, a1);
        ratio_scale(T.nue_numu_ratio, nb0, nb1, b0, b1);
        const double idx_scale = pow(e / 24.0900951261, T.delta_index);
        a0 *= idx_scale; a1 *= idx_scale; b0 *= idx_scale; b1 *= idx_scale;
        // nu/nubar ratio (:178-190)
        double e_nu, e_nb, m_nu, m_nb;
        ratio_scale(T.nu_nubar_ratio, a0, b0, e_nu, e_nb);
        ratio_scale(T.nu_nubar_ratio, a1, b1, m_nu, m_nb);
        double o0 = nubar < 0 ? e_nb : e_nu, o1 = nubar < 0 ? m_nb : m_nu;
        // Barr nu/nubar (ModFlux with unit shape parameters, :42-81,106-113)
        const double l10 = log10(e);
        const double cut_e = exp(-e / 650.), cut_mu = exp(-e / 1000.);
        const double gauss36 = exp(-cz * cz * inv_2s36);
        {
            const double A_ave = loglog(T.ave_e, l10), A_shape = loglog(T.shape_e, l10) * cut_e;
            const double mod = T.nubar_sys * (A_ave - (1.5 * (A_shape * inv_norm36 * gauss36) - 0.7 * A_shape));
            o0 *= nubar < 0 ? fmax(0., 1. / (1 + 0.5 * mod)) : fmax(0., 1. + 0.5 * mod);
        }
        {
            const double A_ave = loglog(T.ave_mu, l10), A_shape = 2.5 * (loglog(T.shape_mu, l10) * cut_mu);
            const double mod = T.nubar_sys * (A_ave - ((A_shape * inv_norm36 * gauss36) - 0.6 * A_shape));
            o1 *= nubar < 0 ? fmax(0., 1. / (1 + 0.5 * mod)) : fmax(0., 1. + 0.5 * mod);
        }
        // Barr up/horizontal: nue only (:83-104)
        {
            const double A_shape = fabs(T.uphor) * (loglog(T.uphor_e, l10) * cut_e);
            const double s = T.uphor == 0 ? 0.0 : (T.uphor > 0 ? 1.0 : -1.0);
            o0 *= 1 - 0.3 * s * norm_fcn(cz, A_shape, inv_norm35, inv_2s35);
        }
        out[2 * i] = (IO)o0;
        out[2 * i + 1] = (IO)o1;
    }
}

// ---- fit-loop form: parameter-independent terms once, cheap per-template apply -----------------
// Everything transcendental in apply_sys_kernel depends on the event only, not on the five systematic
// parameters:   t0 = ln(E / E0)                       (spectral index: (E/E0)^delta = exp(delta t0))
//               t1 = ModFlux(nue; E, cz), t2 = ModFlux(numu; E, cz)    (Barr nu/nubar, unit shape parameters)
//               t3 = LogLog(E) exp(-E/650) N(cz; 0.35)                 (Barr up/horizontal: 1 - 0.3 uphor t3)
// flux_barr_terms_kernel stores them once per container ([n,4] doubles); flux_barr_apply_kernel then needs one
// exp, nine reciprocals and ~30 FMAs per event and is HBM-bound (80 B/event) instead of bound by ~12 FP64
// transcendentals.  ratio_scale is homogeneous of degree one, so the spectral-index factor is applied last.
template <typename IO>
__global__ void __launch_bounds__(256)
flux_barr_terms_kernel(const __grid_constant__ BarrTable T, const IO *__restrict__ energy,
                       const IO *__restrict__ coszen, int64_t n, double *__restrict__ terms) {
    const double inv_norm36 = 1.0 / sqrt(2 * M_PI * 0.36 * 0.36), inv_2s36 = 1.0 / (2 * 0.36 * 0.36);
    const double inv_norm35 = 1.0 / sqrt(2 * M_PI * 0.35 * 0.35), inv_2s35 = 1.0 / (2 * 0.35 * 0.35);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double e = (double)__ldg(energy + i), cz = (double)__ldg(coszen + i);
        const double l10 = log10(e);
        const double cut_e = exp(-e / 650.), cut_mu = exp(-e / 1000.);
        const double gauss36 = exp(-cz * cz * inv_2s36), gauss35 = exp(-cz * cz * inv_2s35);
        const double As_e = loglog(T.shape_e, l10) * cut_e, As_mu = 2.5 * (loglog(T.shape_mu, l10) * cut_mu);
        double4 t;
        t.x = log(e / 24.0900951261);
        t.y = loglog(T.ave_e, l10) - (1.5 * (As_e * inv_norm36 * gauss36) - 0.7 * As_e);
        t.z = loglog(T.ave_mu, l10) - ((As_mu * inv_norm36 * gauss36) - 0.6 * As_mu);
        t.w = loglog(T.uphor_e, l10) * cut_e * inv_norm35 * gauss35;
        reinterpret_cast<double4 *>(terms)[i] = t;
    }
}

template <typename IO>
__device__ __forceinline__ void flux_barr_apply_range(const BarrTable &T, const double *__restrict__ terms,
                                                      const IO *__restrict__ nu_nom, const IO *__restrict__ nubar_nom,
                                                      int nubar, int64_t n, IO *__restrict__ out, int64_t first,
                                                      int64_t stride) {
    const BarrSys S = {T.nue_numu_ratio, T.nu_nubar_ratio, T.delta_index, T.uphor, T.nubar_sys};
    for (int64_t i = first; i < n; i += stride) {
        double o0, o1;
        barr_apply_event<IO>(S, terms, nu_nom, nubar_nom, nubar, i, o0, o1);
        out[2 * i] = (IO)o0;
        out[2 * i + 1] = (IO)o1;
    }
}

template <typename IO>
__global__ void __launch_bounds__(256)
flux_barr_apply_kernel(const __grid_constant__ BarrTable T, const double *__restrict__ terms,
                       const IO *__restrict__ nu_nom, const IO *__restrict__ nubar_nom, int nubar, int64_t n,
                       IO *__restrict__ out) {
    flux_barr_apply_range<IO>(T, terms, nu_nom, nubar_nom, nubar, n, out, (int64_t)blockIdx.x * blockDim.x + threadIdx.x,
                              (int64_t)gridDim.x * blockDim.x);
}

// all flavour containers of a template in ONE launch (a fit that floats the flux systematics re-evaluates nu_flux for
// every hypothesis: 12 launches per template dominate an analysis-size sample): block -> (container, rank)
struct FluxBatch {
    int32_t n_containers, blocks_per_container;
    struct Item { const double *terms; const void *nu_nom, *nubar_nom; void *out; int64_t n; int32_t nubar, pad; } c[PISAB_MAX_BATCH];
};
template <typename IO>
__global__ void __launch_bounds__(256)
flux_barr_apply_batch_kernel(const __grid_constant__ BarrTable T, const __grid_constant__ FluxBatch B) {
    const int ci = blockIdx.x / B.blocks_per_container, r = blockIdx.x - ci * B.blocks_per_container;
    const FluxBatch::Item &C = B.c[ci];
    flux_barr_apply_range<IO>(T, C.terms, (const IO *)C.nu_nom, (const IO *)C.nubar_nom, C.nubar, C.n, (IO *)C.out,
                              (int64_t)r * blockDim.x + threadIdx.x, (int64_t)B.blocks_per_container * blockDim.x);
}

static void fill_barr_table(BarrTable &T, double nue_numu_ratio, double nu_nubar_ratio, double delta_index,
                            double uphor, double nubar_sys) {
    // constants of ModFlux / modRatioUpHor (barr_parameterization.py:44-61,85-94)
    const double e1max_mu = 3., e2max_mu = 43, e1max_e = 2.5, e2max_e = 10, x1e = 0.5, x2e = 3.;
    const double z1max_mu = 0.6, z2max_mu = 5., z1max_e = 0.3, z2max_e = 5., x1z = 0.5, x2z = 2.;
    T.ave_mu = fold_loglog(e1max_mu, e2max_mu, x1e, x2e);
    T.shape_mu = fold_loglog(z1max_mu, z2max_mu, x1z, x2z);
    T.ave_e = fold_loglog(e1max_mu + e1max_e, e2max_mu + e2max_e, x1e, x2e);
    T.shape_e = fold_loglog(z1max_mu + z1max_e, z2max_mu + z2max_e, x1z, x2z);
    T.uphor_e = fold_loglog(z1max_e + z1max_mu, z2max_e + z2max_mu, x1z, x2z);
    T.nue_numu_ratio = nue_numu_ratio; T.nu_nubar_ratio = nu_nubar_ratio; T.delta_index = delta_index;
    T.uphor = uphor; T.nubar_sys = nubar_sys;
}

static int flux_grid(int64_t n) {
    const int sms = sm_count() > 0 ? sm_count() : 148;
    const int64_t want = (n + 255) / 256;
    return (int)(want < (int64_t)sms * 8 ? (want < 1 ? 1 : want) : (int64_t)sms * 8);
}

template <typename IO>
static int flux_terms_impl(const IO *d_energy, const IO *d_coszen, int64_t n, double *d_terms, void *stream) {
    if (n < 0 || (n > 0 && (!d_energy || !d_coszen || !d_terms))) { set_error("bad event arrays"); return PISAB_ERR_ARG; }
    if ((uintptr_t)d_terms % 32 != 0) { set_error("d_terms must be 32-byte aligned"); return PISAB_ERR_ARG; }
    if (n == 0) return PISAB_OK;
    BarrTable T;
    fill_barr_table(T, 1, 1, 0, 0, 0);
    flux_barr_terms_kernel<IO><<<flux_grid(n), 256, 0, (cudaStream_t)stream>>>(T, d_energy, d_coszen, n, d_terms);
    note_launch();
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}

template <typename IO>
static int flux_apply_impl(const double *d_terms, const IO *d_nu, const IO *d_nubar, int32_t nubar,
                           double nue_numu_ratio, double nu_nubar_ratio, double delta_index, double uphor,
                           double nubar_sys, int64_t n, IO *d_out, void *stream) {
    if (n < 0 || (n > 0 && (!d_terms || !d_nu || !d_nubar || !d_out))) { set_error("bad event arrays"); return PISAB_ERR_ARG; }
    if (nubar != 1 && nubar != -1) { set_error("nubar must be +1 or -1"); return PISAB_ERR_ARG; }
    if ((uintptr_t)d_terms % 32 != 0) { set_error("d_terms must be 32-byte aligned"); return PISAB_ERR_ARG; }
    if ((uintptr_t)d_nu % (2 * sizeof(IO)) != 0 || (uintptr_t)d_nubar % (2 * sizeof(IO)) != 0) { set_error("nominal flux rows must be aligned to their size"); return PISAB_ERR_ARG; }
    if (n == 0) return PISAB_OK;
    BarrTable T;
    fill_barr_table(T, nue_numu_ratio, nu_nubar_ratio, delta_index, uphor, nubar_sys);
    flux_barr_apply_kernel<IO><<<flux_grid(n), 256, 0, (cudaStream_t)stream>>>(T, d_terms, d_nu, d_nubar, nubar, n, d_out);
    note_launch();
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}

template <typename IO>
static int flux_apply_batch_impl(const pisab_flux_item_t *items, int32_t n_items, double nue_numu_ratio,
                                 double nu_nubar_ratio, double delta_index, double uphor, double nubar_sys, void *stream) {
    if (!items || n_items < 1 || n_items > PISAB_MAX_BATCH) { set_error("flux batch: 1..%d containers", PISAB_MAX_BATCH); return PISAB_ERR_ARG; }
    FluxBatch B = {};
    B.n_containers = n_items;
    int64_t n_max = 0;
    for (int c = 0; c < n_items; ++c) {
        const pisab_flux_item_t &S = items[c];
        if (S.n < 0 || (S.n > 0 && (!S.d_terms || !S.d_nu_flux_nominal || !S.d_nubar_flux_nominal || !S.d_nu_flux))) { set_error("flux batch: container %d: bad event arrays", c); return PISAB_ERR_ARG; }
        if (S.nubar != 1 && S.nubar != -1) { set_error("nubar must be +1 or -1"); return PISAB_ERR_ARG; }
        if ((uintptr_t)S.d_terms % 32 != 0) { set_error("d_terms must be 32-byte aligned"); return PISAB_ERR_ARG; }
        if ((uintptr_t)S.d_nu_flux_nominal % (2 * sizeof(IO)) != 0 || (uintptr_t)S.d_nubar_flux_nominal % (2 * sizeof(IO)) != 0) { set_error("nominal flux rows must be aligned to their size"); return PISAB_ERR_ARG; }
        B.c[c].terms = S.d_terms; B.c[c].nu_nom = S.d_nu_flux_nominal; B.c[c].nubar_nom = S.d_nubar_flux_nominal;
        B.c[c].out = S.d_nu_flux; B.c[c].n = S.n; B.c[c].nubar = S.nubar;
        if (S.n > n_max) n_max = S.n;
    }
    if (n_max == 0) return PISAB_OK;
    const int sms = sm_count() > 0 ? sm_count() : 148;
    int64_t bpc = (n_max + 255) / 256;
    const int64_t cap = ((int64_t)sms * 8 + n_items - 1) / n_items;
    if (bpc > cap) bpc = cap;
    B.blocks_per_container = (int32_t)bpc;
    BarrTable T;
    fill_barr_table(T, nue_numu_ratio, nu_nubar_ratio, delta_index, uphor, nubar_sys);
    flux_barr_apply_batch_kernel<IO><<<(unsigned)(bpc * n_items), 256, 0, (cudaStream_t)stream>>>(T, B);
    note_launch();
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}

template <typename IO>
static int flux_impl(const IO *d_energy, const IO *d_coszen, const IO *d_nu, const IO *d_nubar, int32_t nubar,
                     double nue_numu_ratio, double nu_nubar_ratio, double delta_index, double uphor,
                     double nubar_sys, int64_t n, IO *d_out, void *stream) {
    if (n < 0 || (n > 0 && (!d_energy || !d_coszen || !d_nu || !d_nubar || !d_out))) { set_error("bad event arrays"); return PISAB_ERR_ARG; }
    if (nubar != 1 && nubar != -1) { set_error("nubar must be +1 or -1"); return PISAB_ERR_ARG; }
    if (n == 0) return PISAB_OK;
    BarrTable T;
    fill_barr_table(T, nue_numu_ratio, nu_nubar_ratio, delta_index, uphor, nubar_sys);
    const int sms = sm_count() > 0 ? sm_count() : 148;
    int64_t want = (n + 255) / 256;
    const int grid = (int)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);
    flux_barr_simple_kernel<IO><<<grid, 256, 0, (cudaStream_t)stream>>>(T, d_energy, d_coszen, d_nu, d_nubar, nubar, n, d_out);
    note_launch();
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}

} // namespace pisab

using namespace pisab;

extern "C" {
int pisab_flux_barr_simple_f64(const double *d_energy, const double *d_coszen, const double *d_nu_flux_nominal,
                               const double *d_nubar_flux_nominal, int32_t nubar, double nue_numu_ratio,
                               double nu_nubar_ratio, double delta_index, double barr_uphor_ratio,
                               double barr_nu_nubar_ratio, int64_t n, double *d_nu_flux, void *stream) {
    return flux_impl<double>(d_energy, d_coszen, d_nu_flux_nominal, d_nubar_flux_nominal, nubar, nue_numu_ratio,
                             nu_nubar_ratio, delta_index, barr_uphor_ratio, barr_nu_nubar_ratio, n, d_nu_flux, stream);
}
int pisab_flux_barr_simple_f32(const float *d_energy, const float *d_coszen, const float *d_nu_flux_nominal,
                               const float *d_nubar_flux_nominal, int32_t nubar, double nue_numu_ratio,
                               double nu_nubar_ratio, double delta_index, double barr_uphor_ratio,
                               double barr_nu_nubar_ratio, int64_t n, float *d_nu_flux, void *stream) {
    return flux_impl<float>(d_energy, d_coszen, d_nu_flux_nominal, d_nubar_flux_nominal, nubar, nue_numu_ratio,
                            nu_nubar_ratio, delta_index, barr_uphor_ratio, barr_nu_nubar_ratio, n, d_nu_flux, stream);
}
int pisab_flux_barr_terms_f64(const double *d_energy, const double *d_coszen, int64_t n, double *d_terms, void *stream) {
    return flux_terms_impl<double>(d_energy, d_coszen, n, d_terms, stream);
}
int pisab_flux_barr_terms_f32(const float *d_energy, const float *d_coszen, int64_t n, double *d_terms, void *stream) {
    return flux_terms_impl<float>(d_energy, d_coszen, n, d_terms, stream);
}
int pisab_flux_barr_apply_f64(const double *d_terms, const double *d_nu_flux_nominal, const double *d_nubar_flux_nominal,
                              int32_t nubar, double nue_numu_ratio, double nu_nubar_ratio, double delta_index,
                              double barr_uphor_ratio, double barr_nu_nubar_ratio, int64_t n, double *d_nu_flux,
                              void *stream) {
    return flux_apply_impl<double>(d_terms, d_nu_flux_nominal, d_nubar_flux_nominal, nubar, nue_numu_ratio, nu_nubar_ratio,
                                   delta_index, barr_uphor_ratio, barr_nu_nubar_ratio, n, d_nu_flux, stream);
}
int pisab_flux_barr_apply_batch_f64(const pisab_flux_item_t *items, int32_t n_items, double nue_numu_ratio,
                                    double nu_nubar_ratio, double delta_index, double barr_uphor_ratio,
                                    double barr_nu_nubar_ratio, void *stream) {
    return flux_apply_batch_impl<double>(items, n_items, nue_numu_ratio, nu_nubar_ratio, delta_index, barr_uphor_ratio,
                                         barr_nu_nubar_ratio, stream);
}
int pisab_flux_barr_apply_batch_f32(const pisab_flux_item_t *items, int32_t n_items, double nue_numu_ratio,
                                    double nu_nubar_ratio, double delta_index, double barr_uphor_ratio,
                                    double barr_nu_nubar_ratio, void *stream) {
    return flux_apply_batch_impl<float>(items, n_items, nue_numu_ratio, nu_nubar_ratio, delta_index, barr_uphor_ratio,
                                        barr_nu_nubar_ratio, stream);
}
int pisab_flux_barr_apply_f32(const double *d_terms, const float *d_nu_flux_nominal, const float *d_nubar_flux_nominal,
                              int32_t nubar, double nue_numu_ratio, double nu_nubar_ratio, double delta_index,
                              double barr_uphor_ratio, double barr_nu_nubar_ratio, int64_t n, float *d_nu_flux,
                              void *stream) {
    return flux_apply_impl<float>(d_terms, d_nu_flux_nominal, d_nubar_flux_nominal, nubar, nue_numu_ratio, nu_nubar_ratio,
                                  delta_index, barr_uphor_ratio, barr_nu_nubar_ratio, n, d_nu_flux, stream);
}
}

// ---------------------------------------------------------------------------------------------
// flux.honda_ip: integral-preserving interpolation of an azimuth-averaged Honda table
// ---------------------------------------------------------------------------------------------
// Reference per event (pisa/utils/flux_weights.py:337-349): 20 x splev(log10 E, energy spline, der=1), cumsum * 0.1,
// splrep through the 21 points, splev(coszen, der=1), / E**enpow -- for each of the four primaries.
//   * splev(.., der=1) is FITPACK's splder: a quadratic in s = log10(E) - t(l) on the knot interval l found by
//     its search rule (l clamped to [k1, nk1]: the end polynomials extrapolate);
//   * the per-event spline FIT in coszen is linear in its 21 values, so it is a fixed set of cardinal-spline
//     derivative polynomials, quadratic in u = coszen - break(p) on piece p.
// Both being piecewise quadratic, the flux of a primary on a cell (l, p) is one biquadratic
//     flux = sum_ab K[l][p][primary][a][b] s^a u^b / E**enpow,
// with K expanded once on the host from the reference's own splrep coefficients
// (pisa_b200/utils/flux_weights.py::HondaTable2D._cell_polynomials).  Per event: log10, two interval searches,
// 36 coefficients (288 contiguous bytes, L2-resident table of ~0.5 MB) and 32 FMAs -- instead of 80 spline
// evaluations of 3 coefficients each.
namespace pisab {

constexpr int kHondaMaxKnots = 256, kHondaMaxPieces = 64;

template <typename IO>
__global__ void __launch_bounds__(256, 4)
flux_honda_2d_kernel(const double *__restrict__ knots, int n_knots, const double *__restrict__ cz_breaks,
                     int n_pieces, const double *__restrict__ cells, int enpow, const IO *__restrict__ energy,
                     const IO *__restrict__ coszen, int64_t n, IO *__restrict__ nu_out, IO *__restrict__ nubar_out) {
    __shared__ double s_knots[kHondaMaxKnots], s_breaks[kHondaMaxPieces];
    for (int i = threadIdx.x; i < n_knots; i += blockDim.x) s_knots[i] = knots[i];
    for (int i = threadIdx.x; i < n_pieces; i += blockDim.x) s_breaks[i] = cz_breaks[i];
    __syncthreads();
    const int nk1 = n_knots - 4; // FITPACK's nk1 for k = 3
    const double t0 = s_knots[4], inv_dt = 1.0 / (s_knots[5] - s_knots[4]); // first interior knots
    const double b1 = s_breaks[n_pieces > 1 ? 1 : 0];
    const double inv_db = n_pieces > 2 ? 1.0 / (s_breaks[2] - s_breaks[1]) : 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double e = (double)__ldg(energy + i), cz = (double)__ldg(coszen + i);
        const double x = log10(e);
        // ---- splder: t(l) <= x < t(l+1), l in [k1, nk1] (1-based); uniform interior knots give the guess
        int l = 5 + (int)floor((x - t0) * inv_dt);
        l = l < 4 ? 4 : (l > nk1 ? nk1 : l);
        while (l > 4 && x < s_knots[l - 1]) --l;            // t(l) > x: go down
        while (l < nk1 && !(x < s_knots[l])) ++l;            // x >= t(l+1): go up
        const double s = x - s_knots[l - 1];
        // ---- coszen piece: break(p) <= cz < break(p+1), the last piece closed
        int p = 1 + (int)floor((cz - b1) * inv_db);
        p = p < 0 ? 0 : (p > n_pieces - 1 ? n_pieces - 1 : p);
        while (p > 0 && cz < s_breaks[p]) --p;
        while (p + 1 < n_pieces && cz >= s_breaks[p + 1]) ++p;
        const double u = cz - s_breaks[p];
        // cell layout [a][b][primary]: one 32-byte group holds the coefficient of s^a u^b for the four primaries, so
        // the Horner scheme in u (inner) and s (outer) runs on four accumulators with 9 x 2 128-bit loads and no
        // 36-entry register array (64 registers -> 4 blocks of 256 threads per SM keep the L2 gathers in flight)
        const double2 *K = reinterpret_cast<const double2 *>(cells + ((size_t)(l - 4) * n_pieces + p) * 36);
        double scale = 1.0;
        for (int j = 0; j < enpow; ++j) scale *= e;
        const double inv_scale = 1.0 / scale;
        double out[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int a = 2; a >= 0; --a) {
            double r[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int b = 2; b >= 0; --b) {
                const double2 lo2 = __ldg(K + (a * 3 + b) * 2), hi2 = __ldg(K + (a * 3 + b) * 2 + 1);
                r[0] = fma(r[0], u, lo2.x);
                r[1] = fma(r[1], u, lo2.y);
                r[2] = fma(r[2], u, hi2.x);
                r[3] = fma(r[3], u, hi2.y);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) out[q] = fma(out[q], s, r[q]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) out[q] *= inv_scale;
        if (sizeof(IO) == 8) {
            reinterpret_cast<double2 *>(nu_out)[i] = make_double2(out[0], out[1]);
            reinterpret_cast<double2 *>(nubar_out)[i] = make_double2(out[2], out[3]);
        } else {
            reinterpret_cast<float2 *>(nu_out)[i] = make_float2((float)out[0], (float)out[1]);
            reinterpret_cast<float2 *>(nubar_out)[i] = make_float2((float)out[2], (float)out[3]);
        }
    }
}

template <typename IO>
static int honda_impl(const double *d_knots, int32_t n_knots, const double *d_cz_breaks, int32_t n_pieces,
                      const double *d_cells, int32_t enpow, const IO *d_energy, const IO *d_coszen, int64_t n,
                      IO *d_nu, IO *d_nubar, void *stream) {
    if (!d_knots || !d_cz_breaks || !d_cells || n_knots < 9 || n_knots > kHondaMaxKnots || n_pieces < 1 ||
        n_pieces > kHondaMaxPieces || enpow < 0) {
        set_error("bad flux table");
        return PISAB_ERR_ARG;
    }
    if (n < 0 || (n > 0 && (!d_energy || !d_coszen || !d_nu || !d_nubar))) { set_error("bad event arrays"); return PISAB_ERR_ARG; }
    if ((((uintptr_t)d_nu | (uintptr_t)d_nubar) & (2 * sizeof(IO) - 1)) != 0) { set_error("flux outputs must be aligned to a (nue, numu) pair"); return PISAB_ERR_ARG; }
    if (n == 0) return PISAB_OK;
    const int sms = sm_count() > 0 ? sm_count() : 148;
    int64_t want = (n + 255) / 256;
    const int grid = (int)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);
    flux_honda_2d_kernel<IO><<<grid, 256, 0, (cudaStream_t)stream>>>(d_knots, n_knots, d_cz_breaks, n_pieces, d_cells,
                                                                     enpow, d_energy, d_coszen, n, d_nu, d_nubar);
    note_launch();
    PISAB_CUDA_CHECK(cudaGetLastError());
    return PISAB_OK;
}

} // namespace pisab

extern "C" {
int pisab_flux_honda_2d_f64(const double *d_knots, int32_t n_knots, const double *d_cz_breaks, int32_t n_pieces,
                            const double *d_cells, int32_t enpow, const double *d_energy, const double *d_coszen,
                            int64_t n, double *d_nu_flux_nominal, double *d_nubar_flux_nominal, void *stream) {
    return pisab::honda_impl<double>(d_knots, n_knots, d_cz_breaks, n_pieces, d_cells, enpow, d_energy, d_coszen, n,
                                     d_nu_flux_nominal, d_nubar_flux_nominal, stream);
}
int pisab_flux_honda_2d_f32(const double *d_knots, int32_t n_knots, const double *d_cz_breaks, int32_t n_pieces,
                            const double *d_cells, int32_t enpow, const float *d_energy, const float *d_coszen,
                            int64_t n, float *d_nu_flux_nominal, float *d_nubar_flux_nominal, void *stream) {
    return pisab::honda_impl<float>(d_knots, n_knots, d_cz_breaks, n_pieces, d_cells, enpow, d_energy, d_coszen, n,
                                    d_nu_flux_nominal, d_nubar_flux_nominal, stream);
}
}
