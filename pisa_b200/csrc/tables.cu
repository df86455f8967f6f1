// tables.cu -- host side: pisab_osc_consts_t / pisab_earth_t -> the per-launch device tables.
// No device code in this file (it is also compiled by g++ in scratch/hostemu for the device-math
// emulation harness used during development).
#include <math.h>
#include <string.h>

#include "common.cuh"

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
    // decay_flag == 1 selects the decay branch (build_decay_table, prob3_decay.cuh); like the reference
    // (numba_osc_kernels.py:445-457) every other value means standard oscillations
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
    // vacuum shortcut tables: projectors on the 2nd and 3rd mass eigenstate
    double P[2][3][3][2];
    for (int k = 1; k < 3; ++k)
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                const double ar = U[i][k][0], ai = U[i][k][1], br = U[j][k][0], bi = -U[j][k][1];
                P[k - 1][i][j][0] = ar * br - ai * bi;
                P[k - 1][i][j][1] = ar * bi + ai * br;
            }
    pack_herm(P[0], 1.0, &out->pr2);
    pack_herm(P[1], 1.0, &out->pr3);
    out->hdm21 = 0.5 * d[1];
    out->hdm31 = 0.5 * d[2];
    bool lr_zero = true;
    for (int i = 0; i < 9; ++i) lr_zero = lr_zero && c->lri_pot[i] == 0.0;
    // (with decay the atmosphere is not a pure phase rotation of the mass states: no shortcut; see build_decay_table)
    out->vac_ok = (lr_zero && c->decay_flag != 1) ? 1.0 : 0.0;
    {
        const Herm3 &v = out->vm;
        out->std_matter = (v.d1 == 0.0 && v.d2 == 0.0 && v.r01 == 0.0 && v.i01 == 0.0 && v.r02 == 0.0 &&
                           v.i02 == 0.0 && v.r12 == 0.0 && v.i12 == 0.0) ? 1.0 : 0.0;
    }
    return PISAB_OK;
}

// H_decay = U mat_decay U^dagger (get_H_decay, numba_osc_kernels.py:571-603), halved like H_vac (one_over_two_e, :443-448).
// Antineutrinos: the reference propagates conj(U) mat_decay conj(U)^dagger with conj(U), -a conj(V), -lri; as for the
// standard tables (common.cuh) the same probabilities come from the neutrino code with -hv/E + rho vm + lr and the decay
// term U (-conj(mat_decay)) U^dagger / 2E.
int build_decay_table(const pisab_osc_consts_t *c, DecayTable *out) {
    if (!c || !out) { set_error("null osc consts"); return PISAB_ERR_ARG; }
    double U[3][3][2], G[3][3][2];
    memcpy(U, c->mix, sizeof U);
    memcpy(G, c->mat_decay, sizeof G);
    for (int which = 0; which < 2; ++which)
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double re = 0, im = 0;
                for (int k = 0; k < 3; ++k)
                    for (int l = 0; l < 3; ++l) {
                        // U[i][k] * g[k][l] * conj(U[j][l]),  g = mat_decay or -conj(mat_decay)
                        const double gr = which ? -G[k][l][0] : G[k][l][0], gi = G[k][l][1];
                        const double ar = U[i][k][0], ai = U[i][k][1], br = U[j][l][0], bi = -U[j][l][1];
                        const double tr = ar * gr - ai * gi, ti = ar * gi + ai * gr;
                        re += tr * br - ti * bi;
                        im += tr * bi + ti * br;
                    }
                out->hd[which][i][j][0] = 0.5 * re;
                out->hd[which][i][j][1] = 0.5 * im;
            }
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

} // namespace pisab
