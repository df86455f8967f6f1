/*
 * oracle/layers_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's Earth-layer geometry
 * (icecube/pisa, pisa/stages/osc/layers.py:38-169 extCalcLayers and
 * :308-335 computeMinLengthToLayers, :411-439 weight_density_to_YeFrac).
 * Checker only; see prob3_oracle.c for the rules on who may call it.
 *
 * Parity status: PINNED against the closed-form numbers in the reference's
 * own test_layers_1..4 (layers.py:485-772) and against extCalcLayers outputs
 * generated in the build container (tests/golden/ref_layers_*.npz).
 *
 * All arithmetic is IEEE double/float with the reference's operation order
 * (numpy evaluates `r**2.*cz**2. - r**2. + radii**2.` left to right, x**2 as
 * x*x); compile with -ffp-contract=off.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#ifdef ORACLE_F32
typedef float real_t;
#define R_SQRT sqrtf
#define SUFFIX(name) name##_f32
#else
typedef double real_t;
#define R_SQRT sqrt
#define SUFFIX(name) name##_f64
#endif

#define MAX_RADII 64

/* layers.py:308-335 computeMinLengthToLayers (float64 python scalars, cast to FTYPE at :335) */
void SUFFIX(oracle_coszen_limits)(const real_t *radii, int n_radii, double r_detector,
                                  real_t *coszen_limit) {
    for (int i = 0; i < n_radii; ++i) {
        double rad = (double)radii[i];
        double x = rad >= r_detector ? 1.0 : -sqrt(1 - ((rad * rad) / (r_detector * r_detector)));
        coszen_limit[i] = (real_t)x;
    }
}

/*
 * layers.py:38-169 extCalcLayers for an array of coszen.
 * radii / rhos / coszen_limit are ordered surface (atmosphere shell) first.
 * Outputs are [n_cz, max_layers], zero padded; n_layers[i] = number of
 * segments with length > 0 (:161).  Returns -1 if the reference itself would
 * fail for this geometry (segment/density count mismatch, idx != 2 in the
 * two-root branch).
 *
 * NOTE on dtype: extCalcLayers is numba-jitted; r_detector is a python float
 * (float64) and numba promotes float32 (op) float64 -> float64, so in FP32 mode
 * every root / distance is evaluated in float64 from the float32-valued inputs
 * and only rounded to FTYPE when stored into the output arrays (:77-79,165-167).
 * Comparisons between coszen and coszen_limit are float32 vs float32 (exact
 * under promotion).  The oracle therefore computes in double in both modes.
 */
int SUFFIX(oracle_calc_layers)(const real_t *cz, int64_t n_cz, double r_detector_d,
                               const real_t *rhos, const real_t *coszen_limit, const real_t *radii,
                               int n_radii, int max_layers, real_t *densities, real_t *distances,
                               real_t *n_layers) {
    if (n_radii > MAX_RADII || max_layers < 2 * n_radii) return -2;
    const double r_det = r_detector_d;
    int idx = -1;
    for (int j = 0; j < n_radii; ++j)
        if ((double)radii[j] < r_det) { idx = j; break; }
    if (idx < 0) return -3;

    for (int64_t i = 0; i < n_cz; ++i) {
        const double coszen = (double)cz[i];
        real_t *den = densities + (size_t)i * max_layers;
        real_t *dis = distances + (size_t)i * max_layers;
        for (int j = 0; j < max_layers; ++j) { den[j] = 0; dis[j] = 0; }
        double seg[2 * MAX_RADII + 2];
        real_t rho_out[2 * MAX_RADII + 2];
        int n_seg = 0;

        if (coszen >= (double)coszen_limit[idx]) {
            /* no tangent: only the shells outside the detector are crossed, once (:94-103) */
            double cum[MAX_RADII];
            for (int j = 0; j < idx; ++j)
                cum[j] = -r_det * coszen +
                         sqrt(r_det * r_det * (coszen * coszen) - r_det * r_det + (double)radii[j] * (double)radii[j]);
            /* diff of [0, cum[idx-1], ..., cum[0]] reversed */
            for (int j = 0; j < idx; ++j) {
                double inner = (j == idx - 1) ? 0.0 : cum[j + 1];
                seg[j] = cum[j] - inner;
            }
            for (int j = idx; j < n_radii; ++j) seg[j] = 0;
            n_seg = n_radii;
            for (int j = 0; j < n_radii; ++j) rho_out[j] = rhos[j] * (seg[j] > 0 ? (real_t)1 : (real_t)0);
        } else {
            /* two-root branch (:105-159) */
            double small_roots[MAX_RADII], large_roots[MAX_RADII];
            int n_small = 0, n_large = 0;
            double full[2 * MAX_RADII + 2];
            int n_full = 0;
            for (int j = 0; j < n_radii; ++j) {
                int calc_small = (coszen < (double)coszen_limit[j]) && (coszen_limit[j] <= coszen_limit[idx]);
                int calc_large = (double)coszen_limit[j] > coszen;
                /* `coszen**2` (int exponent) stays in FTYPE, `coszen**2.` in the other branch is float64 */
                double cz2 = (double)((real_t)cz[i] * (real_t)cz[i]);
                double root = sqrt(r_det * r_det * cz2 - r_det * r_det + (double)radii[j] * (double)radii[j]);
                double s = -r_det * coszen * (double)calc_small - root;
                double l = -r_det * coszen * (double)calc_large + root;
                if (s > 0) small_roots[n_small++] = s; /* nan > 0 is False */
                large_roots[j] = l;
            }
            full[n_full++] = 0;
            for (int j = 0; j < n_small; ++j) full[n_full++] = small_roots[j];
            for (int j = n_radii - 1; j >= 0; --j)
                if (large_roots[j] > 0) { full[n_full++] = large_roots[j]; ++n_large; }
            n_seg = n_full - 1;
            for (int j = 0; j < n_seg; ++j) seg[n_seg - 1 - j] = full[j + 1] - full[j];
            /* densities: crossed shells outer->inner, then inner->outer without the
             * innermost and the outermost (:148-155) */
            int crossed[MAX_RADII], n_crossed = 0;
            for (int j = 0; j < n_radii; ++j)
                if ((double)coszen_limit[j] > coszen) crossed[n_crossed++] = j;
            int n_den = 0;
            for (int j = 0; j < n_crossed; ++j) rho_out[n_den++] = rhos[crossed[j]];
            for (int j = n_crossed - 2; j >= 1; --j) rho_out[n_den++] = rhos[crossed[j]];
            if (n_den != n_seg) return -1; /* numpy would raise on the shape mismatch (:158) */
            for (int j = 0; j < n_seg; ++j) rho_out[j] *= (seg[j] > 0 ? (real_t)1 : (real_t)0);
        }
        real_t cnt = 0;
        for (int j = 0; j < n_seg; ++j) {
            if (seg[j] > 0) cnt += 1;
            if (j < max_layers) { den[j] = rho_out[j]; dis[j] = (real_t)seg[j]; }
        }
        if (n_layers) n_layers[i] = cnt;
    }
    return 0;
}
