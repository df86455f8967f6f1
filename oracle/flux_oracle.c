/*
 * oracle/flux_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's Barr-style flux systematics
 * (icecube/pisa, SURVEY.md 8f.1 "next" row):
 *   - pisa/stages/flux/barr_simple.py:107-197  apply_ratio_scale, spectral_index_scale,
 *                                              apply_sys_kernel
 *   - pisa/utils/barr_parameterization.py:17-113  sign, LogLogParam, norm_fcn, ModFlux,
 *                                              modRatioUpHor, modRatioNuBar
 * FP64 only (numba promotes the float literals to float64, see prob3_oracle.c).
 * Checker only; see prob3_oracle.c for the rules on who may call it.
 * Parity status: PINNED against outputs of the unmodified reference's
 * apply_sys_vectorized (tests/golden/ref_flux_f8.npz, made by make_golden.py).
 */
#include <math.h>
#include <stdint.h>

/* barr_parameterization.py:17-24 */
static double sgn(double v) { return v == 0 ? 0.0 : (v >= 0 ? 1.0 : -1.0); }

/* barr_parameterization.py:26-35 */
static double loglog_param(double e, double y1, double y2, double x1, double x2, int use_cutoff, double cutoff) {
    double nu_nubar = sgn(y2);
    y1 = sgn(y1) * log10(fabs(y1) + 0.0001);
    y2 = log10(fabs(y2 + 0.0001));
    double mod = nu_nubar * pow(10., (((y2 - y1) / (x2 - x1)) * (log10(e) - x1) + y1 - 2.));
    if (use_cutoff) mod *= exp(-1. * e / cutoff);
    return mod;
}

/* barr_parameterization.py:37-40 */
static double norm_fcn(double x, double A, double sigma) {
    return A / sqrt(2 * M_PI * pow(sigma, 2)) * exp(-pow(x, 2) / (2 * pow(sigma, 2)));
}

/* barr_parameterization.py:42-81, called with all eight shape parameters = 1 (:108) */
static double mod_flux(int flav, double e, double cz) {
    const double e1max_mu = 3., e2max_mu = 43, e1max_e = 2.5, e2max_e = 10, x1e = 0.5, x2e = 3.;
    const double z1max_mu = 0.6, z2max_mu = 5., z1max_e = 0.3, z2max_e = 5.;
    const double nue_cutoff = 650., numu_cutoff = 1000., x1z = 0.5, x2z = 2.;
    if (flav == 1) {
        double A_ave = loglog_param(e, e1max_mu, e2max_mu, x1e, x2e, 0, 0);
        double A_shape = 2.5 * loglog_param(e, z1max_mu, z2max_mu, x1z, x2z, 1, numu_cutoff);
        return A_ave - (norm_fcn(cz, A_shape, 0.36) - 0.6 * A_shape);
    }
    double A_ave = loglog_param(e, e1max_mu + e1max_e, e2max_mu + e2max_e, x1e, x2e, 0, 0);
    double A_shape = 1. * loglog_param(e, z1max_mu + z1max_e, z2max_mu + z2max_e, x1z, x2z, 1, nue_cutoff);
    return A_ave - (1.5 * norm_fcn(cz, A_shape, 0.36) - 0.7 * A_shape);
}

/* barr_parameterization.py:83-104 */
static double mod_ratio_uphor(int flav, double e, double cz, double uphor) {
    const double z1max_mu = 0.6, z2max_mu = 5., z1max_e = 0.3, z2max_e = 5., nue_cutoff = 650., x1z = 0.5, x2z = 2.;
    if (flav == 0) {
        double A_shape = 1. * fabs(uphor) * loglog_param(e, (z1max_e + z1max_mu), (z2max_e + z2max_mu), x1z, x2z, 1, nue_cutoff);
        return 1 - 0.3 * sgn(uphor) * norm_fcn(cz, A_shape, 0.35);
    }
    return 1.;
}

/* barr_parameterization.py:106-113 */
static double mod_ratio_nubar(int64_t nubar, int flav, double e, double cz, double nubar_sys) {
    double modfactor = nubar_sys * mod_flux(flav, e, cz);
    if (nubar < 0) return fmax(0., 1. / (1 + 0.5 * modfactor));
    return fmax(0., 1. + 0.5 * modfactor);
}

/* barr_simple.py:107-136 (sum_constant is always True at the call sites :158-190) */
static void apply_ratio_scale(double ratio_scale, double in1, double in2, double *out) {
    if (in1 == 0. && in2 == 0.) { out[0] = 0.; out[1] = 0.; return; }
    double orig_ratio = in1 / in2;
    double orig_sum = in1 + in2;
    double nw = orig_sum / (1. + ratio_scale * orig_ratio);
    out[0] = ratio_scale * orig_ratio * nw;
    out[1] = nw;
}

/* barr_simple.py:145-197 apply_sys_kernel over n events; nu_flux_nominal / nubar_flux_nominal / out are [n,2] */
void oracle_flux_barr_simple(const double *energy, const double *coszen, const double *nu_nom,
                             const double *nubar_nom, int64_t nubar, double nue_numu_ratio,
                             double nu_nubar_ratio, double delta_index, double barr_uphor_ratio,
                             double barr_nu_nubar_ratio, int64_t n, double *out) {
#pragma omp parallel for
    for (int64_t i = 0; i < n; ++i) {
        double nu[2], nb[2], nue[2], numu[2];
        apply_ratio_scale(nue_numu_ratio, nu_nom[2 * i], nu_nom[2 * i + 1], nu);
        apply_ratio_scale(nue_numu_ratio, nubar_nom[2 * i], nubar_nom[2 * i + 1], nb);
        double idx_scale = pow(energy[i] / 24.0900951261, delta_index); /* :139-142 */
        nu[0] *= idx_scale; nu[1] *= idx_scale; nb[0] *= idx_scale; nb[1] *= idx_scale;
        apply_ratio_scale(nu_nubar_ratio, nu[0], nb[0], nue);
        apply_ratio_scale(nu_nubar_ratio, nu[1], nb[1], numu);
        double o0 = nubar < 0 ? nue[1] : nue[0];
        double o1 = nubar < 0 ? numu[1] : numu[0];
        o0 *= mod_ratio_nubar(nubar, 0, energy[i], coszen[i], barr_nu_nubar_ratio);
        o1 *= mod_ratio_nubar(nubar, 1, energy[i], coszen[i], barr_nu_nubar_ratio);
        o0 *= mod_ratio_uphor(0, energy[i], coszen[i], barr_uphor_ratio);
        o1 *= mod_ratio_uphor(1, energy[i], coszen[i], barr_uphor_ratio);
        out[2 * i] = o0;
        out[2 * i + 1] = o1;
    }
}
