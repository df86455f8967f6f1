// Host emulation of the per-event device math (development tool, see cuda_shim.h).
#ifndef PISAB_HOST_EMU
#define PISAB_HOST_EMU
#endif
#include "../../pisa_b200/csrc/prob3_decay.cuh"
#include "../../pisa_b200/csrc/prob3_walk.cuh"
#include "../../pisa_b200/csrc/tables.cu"
#include <stdarg.h>
namespace pisab {
void set_error(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); }
}
using namespace pisab;
extern "C" int emu_propagate(const pisab_osc_consts_t *c, const pisab_earth_t *e, int nubar, int flav,
                             const double *energy, const double *coszen, int64_t n, double *probability,
                             double *prob_e, double *prob_mu) {
    OscTable ot; EarthTable et;
    int rc = build_osc_table(c, &ot); if (rc) return rc;
    rc = build_earth_table(e, &et); if (rc) return rc;
#pragma omp parallel for
    for (int64_t i = 0; i < n; ++i) {
        const double inv_e = rcp_fast(energy[i]);
        H0Reg h0;
        h0.h = herm_axpy(inv_e, ot.hv[nubar > 0 ? 0 : 1], ot.lr);
        h0.set_poly();
        if (probability) {
            Propagator<3, 3> P;
            propagate_earth<3, 3, false>(h0, ot, et, coszen[i], inv_e, nubar, 0, P);
            for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) probability[i * 9 + a * 3 + b] = P.prob(b, a);
        }
        if (prob_e) {
            Propagator<1, 2> P;
            if (ot.std_matter != 0.0) propagate_earth<1, 2, true>(h0, ot, et, coszen[i], inv_e, nubar, flav, P);
            else propagate_earth<1, 2, false>(h0, ot, et, coszen[i], inv_e, nubar, flav, P);
            prob_e[i] = P.prob(0, 0); prob_mu[i] = P.prob(0, 1);
        }
    }
    return 0;
}

// FP32 mode (prob3_mp.cuh): eigenvalues / phase arguments in FP64, matrices and state in float
extern "C" int emu_propagate_mp(const pisab_osc_consts_t *c, const pisab_earth_t *e, int nubar, int flav,
                                const double *energy, const double *coszen, int64_t n, double *probability,
                                double *prob_e, double *prob_mu) {
    OscTable ot; EarthTable et;
    int rc = build_osc_table(c, &ot); if (rc) return rc;
    rc = build_earth_table(e, &et); if (rc) return rc;
#pragma omp parallel for
    for (int64_t i = 0; i < n; ++i) {
        const double inv_e = rcp_fast(energy[i]);
        const Herm3 hh = herm_axpy(inv_e, ot.hv[nubar > 0 ? 0 : 1], ot.lr);
        if (probability) {
            H0MP<false> h0; h0.init(hh);
            float2 buf[18];
            PropagatorSmemF<3, 3> P{buf, 1};
            propagate_earth<3, 3, false>(h0, ot, et, coszen[i], inv_e, nubar, 0, P);
            for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) probability[i * 9 + a * 3 + b] = P.prob(b, a);
        }
        if (prob_e) {
            float2 buf[9];
            PropagatorSmemF<1, 2> P{buf, 1};
            if (ot.std_matter != 0.0) { H0MP<true> h0; h0.init(hh); propagate_earth<1, 2, true>(h0, ot, et, coszen[i], inv_e, nubar, flav, P); }
            else { H0MP<false> h0; h0.init(hh); propagate_earth<1, 2, false>(h0, ot, et, coszen[i], inv_e, nubar, flav, P); }
            prob_e[i] = P.prob(0, 0); prob_mu[i] = P.prob(0, 1);
        }
    }
    return 0;
}

// FP32 mode, two events per thread (lane-packed float part): events (2k, 2k+1) are propagated together; the caller
// orders the events so that pairs cross the same shells.  mismatch[k] = 1 where a pair disagreed on a decision.
extern "C" int emu_propagate_mp_pairs(const pisab_osc_consts_t *c, const pisab_earth_t *e, int nubar, int flav,
                                      const double *energy, const double *coszen, int64_t n, double *prob_e,
                                      double *prob_mu, int *mismatch) {
    OscTable ot; EarthTable et;
    int rc = build_osc_table(c, &ot); if (rc) return rc;
    rc = build_earth_table(e, &et); if (rc) return rc;
    if (n % 2) return 1;
#pragma omp parallel for
    for (int64_t k = 0; k < n / 2; ++k) {
        const double cz[2] = {coszen[2 * k], coszen[2 * k + 1]};
        const double inv_e[2] = {rcp_fast(energy[2 * k]), rcp_fast(energy[2 * k + 1])};
        const Herm3 ha = herm_axpy(inv_e[0], ot.hv[nubar > 0 ? 0 : 1], ot.lr), hb = herm_axpy(inv_e[1], ot.hv[nubar > 0 ? 0 : 1], ot.lr);
        float4 buf[9];
        PropagatorSmemP<1, 2> P{buf, 1};
        bool bad = false;
        float2 hbuf[15];
        if (ot.std_matter != 0.0) { H0MP2<true> h0; h0.col = hbuf; h0.pitch = 1; h0.init(ha, hb); propagate_earth_pair<1, 2, true>(h0, ot, et, cz, inv_e, nubar, flav, P, bad); }
        else { H0MP2<false> h0; h0.col = hbuf; h0.pitch = 1; h0.init(ha, hb); propagate_earth_pair<1, 2, false>(h0, ot, et, cz, inv_e, nubar, flav, P, bad); }
        const f2 pe = P.prob_r(0, 0), pm = P.prob_r(0, 1);
        prob_e[2 * k] = pe.x; prob_e[2 * k + 1] = pe.y; prob_mu[2 * k] = pm.x; prob_mu[2 * k + 1] = pm.y;
        mismatch[k] = bad ? 1 : 0;
    }
    return 0;
}

// Neutrino decay (prob3_decay.cuh): general-matrix layers through the same Earth walk
extern "C" int emu_propagate_decay(const pisab_osc_consts_t *c, const pisab_earth_t *e, int nubar, int flav,
                                   const double *energy, const double *coszen, int64_t n, double *probability,
                                   double *prob_e, double *prob_mu) {
    OscTable ot; EarthTable et; DecayTable dt;
    int rc = build_osc_table(c, &ot); if (rc) return rc;
    rc = build_earth_table(e, &et); if (rc) return rc;
    rc = build_decay_table(c, &dt); if (rc) return rc;
#pragma omp parallel for
    for (int64_t i = 0; i < n; ++i) {
        const double inv_e = 1.0 / energy[i];
        H0Decay h0;
        h0.init(herm_axpy(nubar > 0 ? inv_e : -inv_e, ot.hv[0], ot.lr), dt, nubar, inv_e);
        if (probability) {
            Propagator<3, 3> P;
            propagate_earth<3, 3, false>(h0, ot, et, coszen[i], inv_e, nubar, 0, P);
            for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) probability[i * 9 + a * 3 + b] = P.prob(b, a);
        }
        if (prob_e) {
            Propagator<1, 2> P;
            propagate_earth<1, 2, false>(h0, ot, et, coszen[i], inv_e, nubar, flav, P);
            prob_e[i] = P.prob(0, 0); prob_mu[i] = P.prob(0, 1);
        }
    }
    return 0;
}
