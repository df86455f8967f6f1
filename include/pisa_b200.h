/*
 * pisa_b200.h -- C ABI of the B200-native oscillation-reweighting + histogramming path.
 *
 * Drop-in boundary for the hot path of icecube/pisa (reference citations are
 * relative to the reference repository root):
 *
 *   osc.prob3 / prob3numba   pisa/stages/osc/prob3.py:329-622,
 *                            pisa/stages/osc/prob3numba/numba_osc_hostfuncs.py:60-70,206-221,
 *                            pisa/stages/osc/prob3numba/numba_osc_kernels.py:121-872
 *   Earth layers             pisa/stages/osc/layers.py:38-169,308-335
 *   utils.hist / histogram   pisa/stages/utils/hist.py:62-218,
 *                            pisa/core/translation.py:90-223 (histogram),
 *                            :228-501 (lookup), :503-597 (find_index)
 *
 * Conventions
 *   - every pointer named d_* is a DEVICE pointer owned by the caller (the Python
 *     Stage classes hold them as torch tensors); struct pointers are HOST pointers.
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream).  Calls only
 *     enqueue work; nothing synchronises unless stated.
 *   - return value 0 = ok; non-zero = PISAB_ERR_*; pisab_last_error() gives the text
 *     (thread-local).  There is no CPU fallback anywhere behind this interface.
 *   - *_f64 / *_f32 select the storage type of event arrays (PISA_FTYPE,
 *     pisa/__init__.py:152-179).  Parameter structs are always double.
 */
#ifndef PISA_B200_H
#define PISA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PISAB_MAX_RADII 64   /* PREM_59layer + atmosphere = 61 shells                     */
#define PISAB_MAX_LAYERS 120 /* layer-matrix cache of the reference, numba_osc_kernels.py:173-177,227; explicit   */
                             /* layer arrays may be up to 2 * PISAB_MAX_RADII wide (PREM_59layer: 122)         */
#define PISAB_MAX_DIMS 4

enum {
    PISAB_OK = 0,
    PISAB_ERR_ARG = 1,         /* bad argument (null pointer, negative size, ...)             */
    PISAB_ERR_CUDA = 2,        /* CUDA runtime error; text in pisab_last_error()              */
    PISAB_ERR_UNSUPPORTED = 3, /* branch of the reference that is out of scope (non-Hermitian */
                               /* matter potential, Earth geometry the reference itself       */
                               /* cannot process) or a combination an entry point does not    */
                               /* fuse (text in pisab_last_error())                           */
    PISAB_ERR_WORKSPACE = 4    /* caller-provided workspace too small                         */
};

/* Argument list of `propagate_array` (numba_osc_hostfuncs.py:60-70), host side.
 * Complex matrices are row-major [3][3][2] = (re, im). */
typedef struct pisab_osc_consts {
    double dm[9];         /* dm[i][j] = m_i^2 - m_j^2 (eV^2), OscParams.dm_matrix            */
    double mix[18];       /* PMNS matrix (un-conjugated; nubar handling is internal)         */
    double mat_pot[18];   /* generalised matter potential / a, diag(1|1.02,0,0) + eps        */
    double mat_decay[18]; /* decay matrix in the mass basis (eV^2), DecayParams.decay_matrix =  */
                          /* diag(0, 0, -i alpha3); read when decay_flag == 1                 */
    double lri_pot[9];    /* long-range-interaction potential (eV), real symmetric            */
    int64_t decay_flag;   /* +1 = oscillations + neutrino decay (numba_osc_kernels.py:445-451, */
                          /* the numpy.linalg.eigvals branch: general-matrix kernels, FP64     */
                          /* arithmetic whatever the storage type; a scan with at least one    */
                          /* such template runs ALL its templates through these kernels);      */
                          /* any other value (-1 in the reference) = standard oscillations     */
} pisab_osc_consts_t;

/* What `Layers` holds after __init__/setElecFrac (layers.py:216-289,308-335):
 * shells ordered surface (atmosphere shell) first. */
typedef struct pisab_earth {
    int32_t n_radii;
    int32_t max_layers;                    /* 2 * n_radii (layers.py:244)                     */
    double r_detector;                     /* r_earth - detector_depth (km)                   */
    double radii[PISAB_MAX_RADII];         /* km, decreasing                                  */
    double rho_e[PISAB_MAX_RADII];         /* electron-fraction weighted densities            */
    double coszen_limit[PISAB_MAX_RADII];  /* tangent direction per shell (1 if r >= r_det)   */
} pisab_earth_t;

/* Regularised output binning as utils.hist builds it (hist.py:86-127): every
 * dimension is either linear-regular in x, linear-regular in log(x), or given by
 * explicit edges (irregular -> searchsorted).  Flat index is row-major. */
enum { PISAB_DIM_LIN = 0, PISAB_DIM_LOG = 1, PISAB_DIM_EDGES = 2 };
typedef struct pisab_binning {
    int32_t n_dims;
    int32_t kind[PISAB_MAX_DIMS];
    int32_t n_bins[PISAB_MAX_DIMS];
    double lo[PISAB_MAX_DIMS];             /* LIN and LOG: the domain (LOG: raw, > 0; the library takes    */
                                           /* its log on the device like the samples', hist.py:118-120)     */
    double hi[PISAB_MAX_DIMS];
    const double *d_edges[PISAB_MAX_DIMS]; /* EDGES: device pointer to n_bins+1 edges         */
} pisab_binning_t;

/* ---- library ------------------------------------------------------------------------ */
const char *pisab_last_error(void);
const char *pisab_version(void);
/* Arithmetic behind the *_f32 propagation entry points (the reference's FP32 mode, PISA_FTYPE=fp32,
 * pisa/__init__.py:152-179, computes everything in float32 / complex64):
 *   PISAB_F32_MATH_MIXED (default)  eigenvalues of the layer Hamiltonians, the phase arguments and the shell geometry
 *                                   in FP64; sin / cos polynomials, divided-difference coefficients, transition
 *                                   matrices, matrix-vector products and the propagation state in float32
 *                                   (csrc/prob3_mp.cuh).  <= 1e-5 absolute on probabilities vs the FP64 result.
 *   PISAB_F32_MATH_FP64             float32 storage only, all arithmetic in FP64 (round-1 behaviour).
 * Process-wide; the *_f64 entry points are not affected. */
enum { PISAB_F32_MATH_FP64 = 0, PISAB_F32_MATH_MIXED = 1 };
int pisab_set_f32_math(int32_t mode);
int pisab_get_f32_math(void);
/* sm count / compute capability of the current device; fails without a CUDA device. */
int pisab_device_info(int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor);

/* ---- Earth layers: Layers.calcLayers / extCalcLayers (layers.py:38-169,339-363) ------- */
/* d_densities, d_distances: [n, earth->max_layers], zero padded, ordered from the
 * production point to the detector; bit-identical to the reference in FP64. */
int pisab_layers_calc_f64(const pisab_earth_t *earth, const double *d_coszen, int64_t n,
                          double *d_densities, double *d_distances, int32_t *d_n_layers,
                          void *stream);
int pisab_layers_calc_f32(const pisab_earth_t *earth, const float *d_coszen, int64_t n,
                          float *d_densities, float *d_distances, int32_t *d_n_layers,
                          void *stream);

/* ---- propagate_array (numba_osc_hostfuncs.py:60-70) ------------------------------------ */
/* Explicit layer arrays, exactly the gufunc's inputs.  nubar: scalar when d_nubar == NULL
 * (aux scalar of a container, prob3.py:582), else per event (+1 / -1, int32).
 * d_probability: [n,3,3], out[i][j] = P(nu_i -> nu_j). */
int pisab_prob3_propagate_layers_f64(const pisab_osc_consts_t *consts, int32_t nubar,
                                     const int32_t *d_nubar, const double *d_energy,
                                     const double *d_densities, const double *d_distances,
                                     int64_t n, int32_t n_layers, double *d_probability,
                                     void *stream);
int pisab_prob3_propagate_layers_f32(const pisab_osc_consts_t *consts, int32_t nubar,
                                     const int32_t *d_nubar, const float *d_energy,
                                     const float *d_densities, const float *d_distances,
                                     int64_t n, int32_t n_layers, float *d_probability,
                                     void *stream);

/* Layers evaluated in-kernel from coszen (prob3.setup_function + compute_function fused,
 * prob3.py:406-409,581-605).  Any of d_probability ([n,3,3]) or (d_prob_e, d_prob_mu)
 * ([n] each, = fill_probs(probability, 0|1, flav), numba_osc_hostfuncs.py:206-221) may be
 * NULL.  flav: scalar when d_flav == NULL.
 * d_order (optional, may be NULL): a permutation of 0..n-1 giving the order in which events are
 * assigned to threads.  Results are always written to the event's own slot, so it changes
 * nothing but speed: pass the events grouped by pisab_layer_count() (number of crossed Earth
 * shells) and every warp walks the same number of layers.  It depends on coszen only, i.e. it
 * is setup-time work like the reference's calcLayers in prob3.setup_function (prob3.py:406-409). */
int pisab_prob3_propagate_earth_f64(const pisab_osc_consts_t *consts, const pisab_earth_t *earth,
                                    int32_t nubar, const int32_t *d_nubar, int32_t flav,
                                    const int32_t *d_flav, const double *d_energy,
                                    const double *d_coszen, const int32_t *d_order, int64_t n,
                                    double *d_probability, double *d_prob_e, double *d_prob_mu,
                                    void *stream);
int pisab_prob3_propagate_earth_f32(const pisab_osc_consts_t *consts, const pisab_earth_t *earth,
                                    int32_t nubar, const int32_t *d_nubar, int32_t flav,
                                    const int32_t *d_flav, const float *d_energy,
                                    const float *d_coszen, const int32_t *d_order, int64_t n,
                                    float *d_probability, float *d_prob_e, float *d_prob_mu,
                                    void *stream);

/* Number of Earth shells crossed per event (count of coszen_limit[j] > coszen, layers.py:112,148):
 * the sort key for d_order above. */
int pisab_layer_count_f64(const pisab_earth_t *earth, const double *d_coszen, int64_t n,
                          int32_t *d_count, void *stream);
int pisab_layer_count_f32(const pisab_earth_t *earth, const float *d_coszen, int64_t n,
                          int32_t *d_count, void *stream);

/* fill_probs (numba_osc_hostfuncs.py:206-221): out[n] = probability[n, initial_flav, flav] */
int pisab_fill_probs_f64(const double *d_probability, int32_t initial_flav, int32_t flav,
                         int64_t n, double *d_out, void *stream);
int pisab_fill_probs_f32(const float *d_probability, int32_t initial_flav, int32_t flav,
                         int64_t n, float *d_out, void *stream);

/* prob3.apply_function (prob3.py:621-622):
 * weights *= nu_flux[:,0]*prob_e + nu_flux[:,1]*prob_mu ; d_nu_flux is [n,2]. */
int pisab_apply_osc_weights_f64(const double *d_nu_flux, const double *d_prob_e,
                                const double *d_prob_mu, int64_t n, double *d_weights,
                                void *stream);
int pisab_apply_osc_weights_f32(const float *d_nu_flux, const float *d_prob_e,
                                const float *d_prob_mu, int64_t n, float *d_weights, void *stream);

/* ---- flux.barr_simple (pisa/stages/flux/barr_simple.py:145-226, utils/barr_parameterization.py) ---- */
/* apply_sys_vectorized: nu_flux[n,2] from the nominal (nue, numu) fluxes of neutrinos and antineutrinos
 * ([n,2] each) and the five Barr-style systematic parameters; nubar = +1 / -1 is the container's aux
 * scalar.  A "next" row of the scope table. */
int pisab_flux_barr_simple_f64(const double *d_energy, const double *d_coszen, const double *d_nu_flux_nominal,
                               const double *d_nubar_flux_nominal, int32_t nubar, double nue_numu_ratio,
                               double nu_nubar_ratio, double delta_index, double barr_uphor_ratio,
                               double barr_nu_nubar_ratio, int64_t n, double *d_nu_flux, void *stream);
int pisab_flux_barr_simple_f32(const float *d_energy, const float *d_coszen, const float *d_nu_flux_nominal,
                               const float *d_nubar_flux_nominal, int32_t nubar, double nue_numu_ratio,
                               double nu_nubar_ratio, double delta_index, double barr_uphor_ratio,
                               double barr_nu_nubar_ratio, int64_t n, float *d_nu_flux, void *stream);

/* Fit-loop form of flux.barr_simple: everything transcendental in apply_sys_kernel depends on the event only.
 * pisab_flux_barr_terms_* stores those four per-event terms once (d_terms: double[n,4], 32-byte aligned);
 * pisab_flux_barr_apply_* then evaluates nu_flux for a set of systematic parameters from the terms and the
 * nominal fluxes (HBM-bound, 80 B/event).  Same result as pisab_flux_barr_simple_* to rounding. */
int pisab_flux_barr_terms_f64(const double *d_energy, const double *d_coszen, int64_t n, double *d_terms, void *stream);
int pisab_flux_barr_terms_f32(const float *d_energy, const float *d_coszen, int64_t n, double *d_terms, void *stream);
int pisab_flux_barr_apply_f64(const double *d_terms, const double *d_nu_flux_nominal, const double *d_nubar_flux_nominal,
                              int32_t nubar, double nue_numu_ratio, double nu_nubar_ratio, double delta_index,
                              double barr_uphor_ratio, double barr_nu_nubar_ratio, int64_t n, double *d_nu_flux,
                              void *stream);
int pisab_flux_barr_apply_f32(const double *d_terms, const float *d_nu_flux_nominal, const float *d_nubar_flux_nominal,
                              int32_t nubar, double nue_numu_ratio, double nu_nubar_ratio, double delta_index,
                              double barr_uphor_ratio, double barr_nu_nubar_ratio, int64_t n, float *d_nu_flux,
                              void *stream);

/* The same for up to PISAB_MAX_BATCH flavour containers in ONE launch (a fit that floats the flux systematics
 * re-evaluates nu_flux for every hypothesis; twelve launches per template dominate an analysis-size sample). */
typedef struct pisab_flux_item {
    const double *d_terms;            /* [n][4] from pisab_flux_barr_terms_*, 32-byte aligned           */
    const void *d_nu_flux_nominal;    /* [n][2] of the call's storage type                                */
    const void *d_nubar_flux_nominal; /* [n][2]                                                           */
    void *d_nu_flux;                  /* [n][2] output                                                    */
    int64_t n;
    int32_t nubar, pad;
} pisab_flux_item_t;
int pisab_flux_barr_apply_batch_f64(const pisab_flux_item_t *items, int32_t n_items, double nue_numu_ratio,
                                    double nu_nubar_ratio, double delta_index, double barr_uphor_ratio,
                                    double barr_nu_nubar_ratio, void *stream);
int pisab_flux_barr_apply_batch_f32(const pisab_flux_item_t *items, int32_t n_items, double nue_numu_ratio,
                                    double nu_nubar_ratio, double delta_index, double barr_uphor_ratio,
                                    double barr_nu_nubar_ratio, void *stream);

/* ---- flux.honda_ip (pisa/stages/flux/honda_ip.py:86-104, pisa/utils/flux_weights.py:267-350) ---------- */
/* calculate_2d_flux_weights for all four primaries of one azimuth-averaged Honda table in one pass:
 * d_nu_flux_nominal[n,2] = (nue, numu), d_nubar_flux_nominal[n,2] = (nuebar, numubar).
 * Tables (all double, built on the host by pisa_b200.utils.flux_weights.HondaTable2D from the reference's own
 * splrep coefficients): d_knots[n_knots] the common knot vector of the 80 energy splines; d_cz_breaks[n_pieces]
 * the break points of the per-event coszen spline fit; d_cells[n_knots-7][n_pieces][3][3][4] the biquadratic
 * sum_ab K[a][b] s^a u^b (s = log10 E - knot, u = coszen - break) that the reference's two-step evaluation
 * reduces to on each (energy interval, coszen piece) cell, index order [a][b][primary], primaries in output
 * order.  16-byte aligned.
 * A "next" row of the scope table. */
int pisab_flux_honda_2d_f64(const double *d_knots, int32_t n_knots, const double *d_cz_breaks, int32_t n_pieces,
                            const double *d_cells, int32_t enpow, const double *d_energy, const double *d_coszen,
                            int64_t n, double *d_nu_flux_nominal, double *d_nubar_flux_nominal, void *stream);
int pisab_flux_honda_2d_f32(const double *d_knots, int32_t n_knots, const double *d_cz_breaks, int32_t n_pieces,
                            const double *d_cells, int32_t enpow, const float *d_energy, const float *d_coszen,
                            int64_t n, float *d_nu_flux_nominal, float *d_nubar_flux_nominal, void *stream);

/* ---- histogramming (hist.py:129-218, translation.py:90-223,417-597) -------------------- */
/* Flat row-major bin index per event, -1 when outside in any dimension.  d_coords[d] are
 * device pointers to the n RAW sample values of dimension d (LOG dims are logged
 * internally exactly like Container.translate, container.py:845-850). Bit-exact vs the
 * reference rule `lo <= x < hi ; (int)((x-lo) * n/(hi-lo))` / searchsorted(right)-1. */
int pisab_hist_index_f64(const pisab_binning_t *binning, const double *const *d_coords,
                         int64_t n, int32_t *d_index, void *stream);
int pisab_hist_index_f32(const pisab_binning_t *binning, const float *const *d_coords, int64_t n,
                         int32_t *d_index, void *stream);

/* Workspace (bytes) the accumulate / fused calls need for n events into n_bins bins. */
int64_t pisab_hist_workspace_bytes(int64_t n, int32_t n_bins);

/* hist[b] = sum w, hist_w2[b] = sum w^2 over events with index b (d_hist_w2 may be NULL).
 * d_weights may be NULL (unweighted counts, hist.py:179-185).  Deterministic: fixed
 * per-warp accumulation order and a fixed-order two-stage reduction (no float atomics)
 * whenever n_bins <= PISAB_DET_MAX_BINS; larger binnings are accumulated EXACTLY in 128-bit fixed point with integer
 * atomics (the sum rounded once), which is bit-reproducible as well.  Output is overwritten, always double. */
#define PISAB_DET_MAX_BINS 1024
int pisab_hist_accumulate_f64(const int32_t *d_index, const double *d_weights, int64_t n,
                              int32_t n_bins, double *d_hist, double *d_hist_w2,
                              void *d_workspace, int64_t workspace_bytes, void *stream);
int pisab_hist_accumulate_f32(const int32_t *d_index, const float *d_weights, int64_t n,
                              int32_t n_bins, double *d_hist, double *d_hist_w2,
                              void *d_workspace, int64_t workspace_bytes, void *stream);

/* lookup / Container.binned_to_array (translation.py:417-501, container.py:981-1012):
 * out[n, width] = flat_hist[index[n], width], 0 outside. */
int pisab_lookup_f64(const int32_t *d_index, const double *d_flat_hist, int64_t n, int32_t width,
                     double *d_out, void *stream);
int pisab_lookup_f32(const int32_t *d_index, const float *d_flat_hist, int64_t n, int32_t width,
                     float *d_out, void *stream);

/* ---- fused template evaluation (prob3 compute+apply and hist in one pass) -- */
/* For every event: probabilities through the Earth (as *_propagate_earth), then
 *   w = weights_in * (nu_flux[0]*prob_e + nu_flux[1]*prob_mu)       (prob3.py:621-622)
 * and hist[index] += w, hist_w2[index] += w*w (hist.py:198-209, error_method 'sumw2').
 * d_weights_out / d_prob_e / d_prob_mu are optional per-event outputs; d_order as above (the
 * histogram is summed in thread order, so a fixed d_order keeps it bit-reproducible). */
int pisab_reweight_hist_f64(const pisab_osc_consts_t *consts, const pisab_earth_t *earth,
                            int32_t nubar, const int32_t *d_nubar, int32_t flav,
                            const int32_t *d_flav, const double *d_energy, const double *d_coszen,
                            const double *d_nu_flux, const double *d_weights_in,
                            const int32_t *d_index, const int32_t *d_order, int64_t n,
                            int32_t n_bins, double *d_hist,
                            double *d_hist_w2, double *d_weights_out, double *d_prob_e,
                            double *d_prob_mu, void *d_workspace, int64_t workspace_bytes,
                            void *stream);
int pisab_reweight_hist_f32(const pisab_osc_consts_t *consts, const pisab_earth_t *earth,
                            int32_t nubar, const int32_t *d_nubar, int32_t flav,
                            const int32_t *d_flav, const float *d_energy, const float *d_coszen,
                            const float *d_nu_flux, const float *d_weights_in,
                            const int32_t *d_index, const int32_t *d_order, int64_t n,
                            int32_t n_bins, double *d_hist,
                            double *d_hist_w2, float *d_weights_out, float *d_prob_e,
                            float *d_prob_mu, void *d_workspace, int64_t workspace_bytes,
                            void *stream);

/* One template = ALL flavour containers of a pipeline in one launch (the fit-loop entry point;
 * SURVEY 8f.1).  Per container: the event arrays as above plus `scale`, the per-container scalar that
 * aeff.aeff multiplies into the weights (livetime * aeff_scale * norms, pisa/stages/aeff/aeff.py:68-88):
 *   w = weights[i] * (nu_flux[i,0]*prob_e + nu_flux[i,1]*prob_mu) * scale
 * d_hist: [n_containers][2][n_bins] (sum w, sum w^2), overwritten.  The descriptor array is a HOST
 * array; all pointers inside are device pointers of the storage type of the entry point.
 * Summation order is fixed by (container sizes, n_containers, device): bit-reproducible run to run; the same
 * container evaluated inside a different batch is split over blocks differently and may differ in the last bits. */
#define PISAB_MAX_BATCH 16
typedef struct pisab_container {
    const void *d_energy, *d_coszen, *d_nu_flux, *d_weights;
    const int32_t *d_index;   /* flat output bin per event, -1 outside                         */
    const int32_t *d_order;   /* optional thread order (see pisab_layer_count_*), may be NULL  */
    void *d_weights_out;      /* optional per-event output weights, may be NULL                */
    int64_t n;
    double scale;
    int32_t nubar, flav;
    int32_t flags;            /* PISAB_CONTAINER_*                                              */
    int32_t pad;
    /* flux.barr_simple evaluated inside the template kernel (PISAB_CONTAINER_FLUX_SYS; all three or none): */
    const double *d_flux_terms;        /* [n][4] from pisab_flux_barr_terms_*, 32-byte aligned          */
    const void *d_nu_flux_nominal;     /* [n][2] of the storage type                                    */
    const void *d_nubar_flux_nominal;  /* [n][2]                                                        */
    /* optional additive per-event term of utils.hist (hist.py:141-145: weights + astro_weights), storage type [n]:
     * w = weights * (flux . prob) * scale + astro_weights.  Not above PISAB_DET_MAX_BINS bins or in a scan. */
    const void *d_astro_weights;
} pisab_container_t;
/* The caller guarantees: n is even and events 2k and 2k+1 cross the same number of Earth shells
 * (pisab_layer_count_*), e.g. because the events are sorted by that count and every class was padded to an even size
 * with zero-weight events.  With float storage and the mixed-precision arithmetic the batched template kernel then
 * handles TWO events per thread in the lanes of the packed FP32 instructions (same results, bit for bit, as one
 * event per thread).  A pair that breaks the promise poisons its weights with NaN instead of histogramming them. */
#define PISAB_CONTAINER_PAIR_ALIGNED 1
/* nu_flux of this container is NOT read: the kernel evaluates flux.barr_simple (barr_simple.py:139-197) per event in
 * registers from d_flux_terms and the two nominal fluxes with the systematics of the call's pisab_flux_sys_t -- the
 * same arithmetic, bit for bit, as pisab_flux_barr_apply_* followed by a template over its output, without the
 * 80 B/event pass that writes nu_flux and the 16 B/event that re-read it.  Not available above PISAB_DET_MAX_BINS
 * bins, with per-event outputs, or in pisab_reweight_hist_scan_* (PISAB_ERR_UNSUPPORTED). */
#define PISAB_CONTAINER_FLUX_SYS 2
typedef struct pisab_flux_sys {   /* parameters of flux.barr_simple, barr_simple.py:41-52 */
    double nue_numu_ratio, nu_nubar_ratio, delta_index, barr_uphor_ratio, barr_nu_nubar_ratio;
} pisab_flux_sys_t;
int64_t pisab_reweight_batch_workspace_bytes(int32_t n_containers, int32_t n_bins);
/* flux_sys: NULL unless a container carries PISAB_CONTAINER_FLUX_SYS. */
int pisab_reweight_hist_batch_f64(const pisab_osc_consts_t *consts, const pisab_earth_t *earth,
                                  const pisab_container_t *containers, int32_t n_containers,
                                  int32_t n_bins, const pisab_flux_sys_t *flux_sys, double *d_hist,
                                  void *d_workspace, int64_t workspace_bytes, void *stream);
int pisab_reweight_hist_batch_f32(const pisab_osc_consts_t *consts, const pisab_earth_t *earth,
                                  const pisab_container_t *containers, int32_t n_containers,
                                  int32_t n_bins, const pisab_flux_sys_t *flux_sys, double *d_hist,
                                  void *d_workspace, int64_t workspace_bytes, void *stream);

/* One hypothesis of a fit in ONE call and two launches: the batched template kernel, then one kernel that
 * reduces the per-block partial histograms, applies the optional per-bin detector-systematics scales of
 * discr_sys.hypersurfaces (pisa/stages/discr_sys/hypersurfaces.py:219-243; d_bin_scales [n_containers][n_bins] or NULL:
 * sum w -> max(s * sum w, 0), sum w^2 -> s^2 sum w^2), sums the containers (MapSet sum, sumw2 errors) and evaluates
 * mod_chi2 (pisa/utils/stats.py:651-695) against d_observed [n_bins] in its last-arriving block.  Outputs: d_hist
 * [n_containers][2][n_bins] (required), d_total [2][n_bins] (optional), d_chi2 one double (optional when d_observed is
 * NULL).  n_bins <= PISAB_DET_MAX_BINS.  Launches of one device must come from one stream at a time. */
int pisab_reweight_hist_chi2_f64(const pisab_osc_consts_t *consts, const pisab_earth_t *earth,
                                 const pisab_container_t *containers, int32_t n_containers, int32_t n_bins,
                                 const pisab_flux_sys_t *flux_sys, const double *d_bin_scales, const double *d_observed, double *d_hist, double *d_total,
                                 double *d_chi2, void *d_workspace, int64_t workspace_bytes, void *stream);
int pisab_reweight_hist_chi2_f32(const pisab_osc_consts_t *consts, const pisab_earth_t *earth,
                                 const pisab_container_t *containers, int32_t n_containers, int32_t n_bins,
                                 const pisab_flux_sys_t *flux_sys, const double *d_bin_scales, const double *d_observed, double *d_hist, double *d_total,
                                 double *d_chi2, void *d_workspace, int64_t workspace_bytes, void *stream);

/* The epilogue of pisab_reweight_hist_chi2_* as a call of its own: sums n_blocks partial histograms per container
 * (d_partials [n_containers][n_blocks][2][n_bins], block order), applies the optional per-bin scales of
 * discr_sys.hypersurfaces (hypersurfaces.py:219-248: weights -> clip(s w, 0), errors -> s errors, i.e. sum w^2 ->
 * s^2 sum w^2), sums the containers and evaluates mod_chi2; outputs as in pisab_reweight_hist_chi2_*.  One launch. */
int pisab_hist_scale_sum_chi2(const double *d_partials, int32_t n_blocks, int32_t n_containers, int32_t n_bins,
                              const double *d_bin_scales, const double *d_observed, double *d_hist, double *d_total,
                              double *d_chi2, void *stream);

/* mod_chi2 (pisa/utils/stats.py:651-695) on device for the scan driver:
 * sum_b (obs-exp)^2 / (sigma^2 + max(exp,1e-10)); result is one double on the device. */
int pisab_mod_chi2(const double *d_expected, const double *d_expected_w2, const double *d_observed,
                   int32_t n_bins, double *d_out, void *stream);

/* Scan / fit driver: sum the container maps of one template (d_hist = [n_containers][2][n_bins] as
 * written by pisab_reweight_hist_batch_*: sum w, sum w^2), take sigma^2 = sum of the sum-w^2 maps
 * (MapSet sum with sumw2 errors, hist.py:205-218) and evaluate mod_chi2 against d_observed.  d_total
 * (optional, [2][n_bins]) receives the summed map and sigma^2; d_out is one double on the device. */
int pisab_template_chi2(const double *d_hist, int32_t n_containers, int32_t n_bins,
                        const double *d_observed, double *d_total, double *d_out, void *stream);

/* Parameter scan (BASELINE configs[4]): n_templates hypotheses in ONE launch.  `consts` is a HOST array of
 * n_templates oscillation-parameter sets; containers as in pisab_reweight_hist_batch_* (d_weights_out is
 * ignored).  d_hist: [n_templates][n_containers][2][n_bins].  pisab_template_chi2_batch then gives one mod_chi2
 * per template (d_out[n_templates]).  For analysis-size samples (1e5 .. 1e6 events) a single template cannot
 * fill the GPU; batching the hypotheses does. */
int64_t pisab_reweight_scan_workspace_bytes(int32_t n_templates, int32_t n_containers, int32_t n_bins, int64_t n_max);
int pisab_reweight_hist_scan_f64(const pisab_osc_consts_t *consts, int32_t n_templates, const pisab_earth_t *earth,
                                 const pisab_container_t *containers, int32_t n_containers, int32_t n_bins,
                                 double *d_hist, void *d_workspace, int64_t workspace_bytes, void *stream);
int pisab_reweight_hist_scan_f32(const pisab_osc_consts_t *consts, int32_t n_templates, const pisab_earth_t *earth,
                                 const pisab_container_t *containers, int32_t n_containers, int32_t n_bins,
                                 double *d_hist, void *d_workspace, int64_t workspace_bytes, void *stream);
int pisab_template_chi2_batch(const double *d_hist, int32_t n_templates, int32_t n_containers, int32_t n_bins,
                              const double *d_observed, double *d_out, void *stream);

/* Any number of bins, static indices: the SORTED plan.  d_perm = stable order of the events by bin index
 * (pisab_sort_order_i32 of the index with out-of-range entries mapped to one key), d_sorted_index = index[perm], both
 * built once.  Per template the kernel reads the plan coalesced, gathers the weights and adds every warp's run of equal
 * bins to the exact 128-bit fixed-point accumulators with one atomic pair: bit-reproducible, independent of grid and
 * order like pisab_hist_accumulate_* above PISAB_DET_MAX_BINS, and ~8x faster than it at 3200 bins.  Workspace:
 * pisab_hist_workspace_bytes(n, n_bins). */
int pisab_hist_accumulate_sorted_f64(const int32_t *d_perm, const int32_t *d_sorted_index, const double *d_weights,
                                     int64_t n, int32_t n_bins, double *d_hist, double *d_hist_w2, void *d_workspace,
                                     int64_t workspace_bytes, void *stream);
int pisab_hist_accumulate_sorted_f32(const int32_t *d_perm, const int32_t *d_sorted_index, const float *d_weights,
                                     int64_t n, int32_t n_bins, double *d_hist, double *d_hist_w2, void *d_workspace,
                                     int64_t workspace_bytes, void *stream);

/* ---- setup-time ordering (stable radix sort of small non-negative integer keys) ---------------------------------
 * d_order[k] = index of the event with the k-th smallest (largest if `descending`) key, ties in input order;
 * d_sorted_keys (optional) = d_keys[d_order].  key_bits: number of significant key bits (0 = 31).  Used once per event
 * sample: grouping events by crossed Earth shells (pisab_layer_count_*; the reference computes its layer arrays once
 * in prob3.setup_function as well, prob3.py:406-409) and the sorted plan of pisab_hist_accumulate_sorted_*. */
int64_t pisab_sort_workspace_bytes(int64_t n);
int pisab_sort_order_i32(const int32_t *d_keys, int64_t n, int32_t key_bits, int32_t descending, int32_t *d_order,
                         int32_t *d_sorted_keys, void *d_workspace, int64_t workspace_bytes, void *stream);

/* ---- planned histogram (the fit-loop form of utils.hist.apply_function) ------------------------------------
 * The bin index of an event is computed once at hist.setup_function (hist.py:86-127) and never changes during a fit.
 * pisab_hist_plan_build turns it, once, into a PLAN: per tile of 2048 events the permutation that groups the tile's
 * events by bin (uint16, stable) and the n_bins + 1 group offsets.  pisab_hist_accumulate_planned_* then produces
 * sum w and sum w^2 from the plan and the current weights: one thread per bin (up to four bins per thread) walks its
 * group in shared memory and accumulates in registers -- no private bins, no atomics, fixed summation order
 * (bit-reproducible), 8 + 2 B/event of DRAM traffic.  Supported for n_bins <= PISAB_DET_MAX_BINS (pisab_hist_plan_bytes
 * returns 0 otherwise: use the sorted plan of pisab_hist_accumulate_sorted_*).  Events whose index is outside [0, n_bins) are dropped, as in pisab_hist_accumulate_*.
 * d_weights must be 16-byte aligned.  Workspace: pisab_hist_workspace_bytes(n, n_bins). */
int64_t pisab_hist_plan_bytes(int64_t n, int32_t n_bins);
int pisab_hist_plan_build(const int32_t *d_index, int64_t n, int32_t n_bins, void *d_plan, int64_t plan_bytes,
                          void *stream);
int pisab_hist_accumulate_planned_f64(const void *d_plan, const double *d_weights, int64_t n, int32_t n_bins,
                                      double *d_hist, double *d_hist_w2, void *d_workspace, int64_t workspace_bytes,
                                      void *stream);
int pisab_hist_accumulate_planned_f32(const void *d_plan, const float *d_weights, int64_t n, int32_t n_bins,
                                      double *d_hist, double *d_hist_w2, void *d_workspace, int64_t workspace_bytes,
                                      void *stream);

/* ---- multi-GPU: the histogram exchange as ONE kernel over NVLink peer memory ------------------
 * One process per GPU.  pisab_exchange_create allocates this rank's exchange buffer and returns its 64-byte CUDA IPC
 * handle; the host side gathers the handles of all ranks (torch.distributed, plumbing) and passes them, in rank order,
 * to pisab_exchange_connect.  pisab_exchange_allreduce then sums d_buf[count] over the ranks IN RANK ORDER, in place,
 * with one kernel launch per rank: peer stores into every rank's slot, system-scope release / acquire flags, fixed-order
 * sum -- the result is bit-identical on every rank and from run to run.  Every rank must make the same sequence of
 * calls; a wait for a peer that never arrives ends after ~4 s and is reported by pisab_exchange_status (0 = ok).
 * pisab_sum_slots is the one-launch rank-ordered sum used after a library all_gather where peer mapping is not
 * available. */
int pisab_exchange_create(int32_t rank, int32_t world, int64_t capacity_doubles, void **ctx_out,
                          unsigned char *handle_out /* [64] */);
int pisab_exchange_connect(void *ctx, const unsigned char *all_handles /* [world][64] */);
int pisab_exchange_allreduce(void *ctx, double *d_buf, int64_t count, void *stream);
int pisab_exchange_status(void *ctx);
int pisab_exchange_disconnect(void *ctx); /* unmap the peers; all ranks, then synchronise the ranks, then destroy */
int pisab_exchange_destroy(void *ctx);
int pisab_sum_slots(const double *d_gathered, int32_t world, int64_t count, double *d_out, void *stream);

/* ---- small stage-API operators ---------------------------------------------------------- */
/* aeff.aeff apply_function (pisa/stages/aeff/aeff.py:68-88): weights[i] *= factor[i] * scale in FTYPE arithmetic
 * (d_factor = weighted_aeff, may be NULL: weights[i] *= scale). */
int pisab_scale_weights_f64(const double *d_factor, double scale, int64_t n, double *d_weights, void *stream);
int pisab_scale_weights_f32(const float *d_factor, double scale, int64_t n, float *d_weights, void *stream);
/* Flat index on the joint binning a + b (utils.hist with a binned calc_mode, hist.py:69-84) from two cached
 * sub-indices: out = a * size_b + b, -1 where either is -1. */
int pisab_joint_index(const int32_t *d_index_a, const int32_t *d_index_b, int32_t size_b, int64_t n,
                      int32_t *d_out, void *stream);
/* utils.hist apply_function with a binned calc_mode (hist.py:131-160): hist = (unc w) @ T, sumw2 = (unc w)^2 @ T,
 * bin_unc2 = (unc^2 w) @ T, T = hist_transform [n_calc][n_out] row-major; d_unc, d_sumw2, d_bin_unc2 may be NULL.
 * Fixed summation order (bit-reproducible). */
int pisab_hist_transform_f64(const double *d_weights, const double *d_unc, const double *d_transform, int32_t n_calc,
                             int32_t n_out, double *d_hist, double *d_sumw2, double *d_bin_unc2, void *stream);
int pisab_hist_transform_f32(const float *d_weights, const float *d_unc, const float *d_transform, int32_t n_calc,
                             int32_t n_out, double *d_hist, double *d_sumw2, double *d_bin_unc2, void *stream);

/* ---- measurement helpers (bench.py) ---------------------------------------------------- */
/* Dependent-chain-free DFMA microbenchmark: runs `iters` x 8 independent FMAs per thread
 * on a full grid and returns achieved FP64 FLOP/s (synchronises). Roofline denominator. */
int pisab_fp64_peak_probe(int32_t iters, double *flops_per_s, double *elapsed_ms);
/* Kernels launched by this library since load / last reset (bench.py's gpu_launches). */
int64_t pisab_launch_count(int32_t reset);
/* Time of the most recent propagation-class kernel measured with CUDA events on its own
 * stream when profiling is enabled (pisab_set_profiling(1)); ms, or negative if none. */
int pisab_set_profiling(int32_t on);
double pisab_last_kernel_ms(void);

#ifdef __cplusplus
}
#endif
#endif /* PISA_B200_H */
