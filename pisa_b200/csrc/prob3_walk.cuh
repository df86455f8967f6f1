// prob3_walk.cuh -- the per-event walk through the Earth's shells, shared by the FP64 and the FP32-mode arithmetic.
#pragma once
#include "prob3_device.cuh"
#include "prob3_mp.cuh"

namespace pisab {

// Per-event propagation through the Earth.  h0 = hv/E + lr (per event), vm scales with rho.
// The path is walked as a sequence of steps (shell, segment length, action) with ONE
// transition_matrix call site, so that the kernel stays small (instruction cache) and a warp of
// events with the same number of crossed shells executes without divergence.
//
//   two-root branch (layers.py:105-159, K crossed shells, first inner shell index 2):
//     shell 0 (atmosphere)       l_0 - l_1            R  = cols(T)
//     shell 1, far side          l_1 - l_2            R <- T R
//     shell 1, near side         s_2 - 0              L  = rows(T)
//     shell j = 2 .. K-2         l_j - l_{j+1}        R <- T R ; L <- L T   (in/out twin, :236-249)
//     shell K-1 (innermost)      l_{K-1} - s_{K-1}    R <- T R
//   no-tangent branch (layers.py:94-103): shells 0 .. idx-1 once each, last one initialises L.
//   Segments of length <= 0 are skipped like in the reference (:233,285).
template <int NR, int NC, bool STD, typename H0, typename PROP>
__device__ __forceinline__ void propagate_earth(const H0 &h0, const OscTable &osc,
                                                const EarthTable &E, double cz, double inv_e,
                                                int nubar, int flav, PROP &P) {
    const Herm3 &vm = osc.vm;
    const double T_SCALE = kTab[18]; // 2 * 2.534: (1/2)(1/hbar c) in GeV/(eV^2 km) (:524), times 2 (M = 2 E lambda)
    enum { ACT_R = 1, ACT_L = 2 };
    const double cz2 = __dmul_rn(cz, cz);
    const double base = __dmul_rn(-E.r_det, cz);
    const int idx = E.idx_first_inner;
    const bool tangent = cz < E.limit[idx];
    bool have_r = false, have_l = false;

    double l_cur = __dadd_rn(base, shell_root_fast(E.rd2, cz2, E.rj2[0])); // large root of current shell
    double sq_cur = 0.0;                                             // sqrt term of current shell
    int j = 0;        // current shell
    int phase = 0;    // 0: walking inwards, 1: near-side piece of the detector shell pending
#ifndef PISAB_NO_VACUUM_SHORTCUT
    // Shell 0 is the atmosphere (rho = 0, layers.py:262-275): with no long-range potential its
    // transition matrix needs no eigenvalue solve.  Every path starts with it (both branches), so it
    // is taken out of the loop; the loop then resumes at shell 1 in exactly the state it would have.
    if (osc.vac_ok != 0.0 && E.rho[0] == 0.0 && idx >= 2) {
        const double sq_next = shell_root_fast(E.rd2, cz2, E.rj2[1]);
        const double l_next = __dadd_rn(base, sq_next);
        const double seg = __dsub_rn(l_cur, l_next);
        l_cur = l_next;
        sq_cur = sq_next;
        j = 1;
        if (seg > 0.0) {
            if constexpr (PROP::kF32) vacuum_columns_mp<NC>(osc, (nubar > 0 ? -T_SCALE : T_SCALE) * seg * inv_e, P);
            else vacuum_columns<NC>(osc, (nubar > 0 ? -T_SCALE : T_SCALE) * seg * inv_e, P);
            have_r = true;
        }
    }
#endif
    for (;;) {
        double seg;
        int act;
        bool last = false;
        const int shell = j;
        if (!tangent) {
            const double l_next = (j + 1 < idx) ? __dadd_rn(base, shell_root_fast(E.rd2, cz2, E.rj2[j + 1])) : 0.0;
            seg = __dsub_rn(l_cur, l_next);
            l_cur = l_next;
            act = (j + 1 == idx) ? ACT_L : ACT_R;
            last = (j + 1 == idx);
            ++j;
        } else if (phase == 1) {
            // near side of the detector shell: small root of shell idx minus 0
            seg = __dsub_rn(base, sq_cur);
            act = ACT_L;
            phase = 0;
            // `shell` is idx-1 here (j was already advanced to idx)
        } else {
            const bool innermost = !(j + 1 < E.n_radii && E.limit[j + 1] > cz);
            if (innermost) {
                seg = __dsub_rn(l_cur, __dsub_rn(base, sq_cur)); // l_j - s_j
                act = ACT_R;
                last = true;
            } else {
                const double sq_next = shell_root_fast(E.rd2, cz2, E.rj2[j + 1]);
                const double l_next = __dadd_rn(base, sq_next);
                seg = __dsub_rn(l_cur, l_next);
                l_cur = l_next;
                sq_cur = sq_next;
                act = (j >= idx) ? (ACT_R | ACT_L) : ACT_R;
                if (j + 1 == idx) phase = 1; // after the far side of shell idx-1 do its near side
                ++j;
            }
        }
        // (Sharing one eigenvalue solve between the two pieces of the detector shell -- same density,
        // different lengths -- was tried: the solve's results have to stay live across the second
        // assembly, which costs 120 B of spills and 9 % of the kernel; not kept.)
        const int rho_shell = (tangent && act == ACT_L) ? idx - 1 : shell;
        if (seg > 0.0) {
            typename PROP::cplx T[3][3];
            if constexpr (is_general_h0<H0>::value) {
                // neutrino decay: non-Hermitian layer Hamiltonian, complex eigenvalues (prob3_decay.cuh)
                h0.layer(E.rho[rho_shell], vm, T_SCALE * seg, T);
            } else if constexpr (PROP::kF32) {
                // FP32 mode: eigenvalues and phase arguments in FP64, everything else in float (prob3_mp.cuh)
                h0.layer(E.rho[rho_shell], vm, T_SCALE * seg, T);
            } else if constexpr (STD) {
                // standard matter: only H[0][0] moves with the density (see H0Reg); H0 and H0^2 are per-event
                const double x = E.rho[rho_shell] * vm.d0;
                double c2, c1, c0;
                h0.poly(x, c2, c1, c0);
                const Eigen eig = eigen_solve(c2, c1, c0);
                assemble_transition<true>(h0.load(), h0.load_sq(), x, eig, T_SCALE * seg, T);
            } else {
                transition_matrix(herm_axpy(E.rho[rho_shell], vm, h0.load()), T_SCALE * seg, T);
            }
            if (act & ACT_R) {
                if (have_r) P.mul_right(T);
                else { P.init_right(T); have_r = true; }
            }
            if (act & ACT_L) {
                if (have_l) P.mul_left(T);
                else { P.init_left(T, flav); have_l = true; }
            }
        }
        if (last) break;
    }
    if (!have_r || !have_l) {
        typename PROP::cplx I[3][3] = {{{1, 0}, {0, 0}, {0, 0}}, {{0, 0}, {1, 0}, {0, 0}}, {{0, 0}, {0, 0}, {1, 0}}};
        if (!have_r) P.init_right(I);
        if (!have_l) P.init_left(I, flav);
    }
}

// The same walk for a PAIR of events handled by one thread in the FP32 mode (prob3_mp.cuh: float part lane-packed).
// Both events must cross the same shells (the host pairs events of equal pisab_layer_count and flags the container
// PISAB_CONTAINER_PAIR_ALIGNED): control flow follows event 0, every decision is re-checked for event 1 and a
// disagreement is reported through `mismatch` (the caller poisons the pair's weights rather than histogram a wrong
// number).  Geometry, densities and segment lengths stay per event and in FP64; a segment of length <= 0 in one lane
// only (exact tangency) becomes t = 0, i.e. the identity for that lane.
template <int NR, int NC, bool STD, typename H0, typename PROP>
__device__ __forceinline__ void propagate_earth_pair(const H0 &h0, const OscTable &osc, const EarthTable &E,
                                                     const double (&cz)[2], const double (&inv_e)[2], int nubar,
                                                     int flav, PROP &P, bool &mismatch) {
    const Herm3 &vm = osc.vm;
    const double T_SCALE = kTab[18];
    enum { ACT_R = 1, ACT_L = 2 };
    const int idx = E.idx_first_inner;
    const bool tangent = cz[0] < E.limit[idx];
    mismatch = (cz[1] < E.limit[idx]) != tangent;
    bool have_r = false, have_l = false;
    double cz2[2], base[2], l_cur[2], sq_cur[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        cz2[e] = __dmul_rn(cz[e], cz[e]);
        base[e] = __dmul_rn(-E.r_det, cz[e]);
        l_cur[e] = __dadd_rn(base[e], shell_root_fast(E.rd2, cz2[e], E.rj2[0]));
        sq_cur[e] = 0.0;
    }
    int j = 0, phase = 0;
#ifndef PISAB_NO_VACUUM_SHORTCUT
    if (osc.vac_ok != 0.0 && E.rho[0] == 0.0 && idx >= 2) {
        double seg[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const double sq_next = shell_root_fast(E.rd2, cz2[e], E.rj2[1]);
            const double l_next = __dadd_rn(base[e], sq_next);
            seg[e] = __dsub_rn(l_cur[e], l_next);
            l_cur[e] = l_next;
            sq_cur[e] = sq_next;
        }
        j = 1;
        if (seg[0] > 0.0 || seg[1] > 0.0) {
            const double sgn = nubar > 0 ? -T_SCALE : T_SCALE;
            vacuum_columns_mp2<NC>(osc, sgn * fmax(seg[0], 0.0) * inv_e[0], sgn * fmax(seg[1], 0.0) * inv_e[1], P);
            have_r = true;
        }
    }
#endif
    for (;;) {
        double seg[2];
        int act;
        bool last = false;
        const int shell = j;
        if (!tangent) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const double l_next = (j + 1 < idx) ? __dadd_rn(base[e], shell_root_fast(E.rd2, cz2[e], E.rj2[j + 1])) : 0.0;
                seg[e] = __dsub_rn(l_cur[e], l_next);
                l_cur[e] = l_next;
            }
            act = (j + 1 == idx) ? ACT_L : ACT_R;
            last = (j + 1 == idx);
            ++j;
        } else if (phase == 1) {
#pragma unroll
            for (int e = 0; e < 2; ++e) seg[e] = __dsub_rn(base[e], sq_cur[e]);
            act = ACT_L;
            phase = 0;
        } else {
            const bool more = j + 1 < E.n_radii;
            const bool innermost = !(more && E.limit[j + 1] > cz[0]);
            mismatch = mismatch || (!(more && E.limit[j + 1] > cz[1]) != innermost);
            if (innermost) {
#pragma unroll
                for (int e = 0; e < 2; ++e) seg[e] = __dsub_rn(l_cur[e], __dsub_rn(base[e], sq_cur[e]));
                act = ACT_R;
                last = true;
            } else {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double sq_next = shell_root_fast(E.rd2, cz2[e], E.rj2[j + 1]);
                    const double l_next = __dadd_rn(base[e], sq_next);
                    seg[e] = __dsub_rn(l_cur[e], l_next);
                    l_cur[e] = l_next;
                    sq_cur[e] = sq_next;
                }
                act = (j >= idx) ? (ACT_R | ACT_L) : ACT_R;
                if (j + 1 == idx) phase = 1;
                ++j;
            }
        }
        const int rho_shell = (tangent && act == ACT_L) ? idx - 1 : shell;
        if (seg[0] > 0.0 || seg[1] > 0.0) {
            typename PROP::cplx T[3][3];
            h0.layer(E.rho[rho_shell], vm, T_SCALE * fmax(seg[0], 0.0), T_SCALE * fmax(seg[1], 0.0), T);
            if (act & ACT_R) {
                if (have_r) P.mul_right(T);
                else { P.init_right(T); have_r = true; }
            }
            if (act & ACT_L) {
                if (have_l) P.mul_left(T);
                else { P.init_left(T, flav); have_l = true; }
            }
        }
        if (last) break;
    }
    if (!have_r || !have_l) {
        typename PROP::cplx I[3][3];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) I[a][b] = typename PROP::cplx{f2{a == b ? 1.0f : 0.0f, a == b ? 1.0f : 0.0f}, f2{0.0f, 0.0f}};
        if (!have_r) P.init_right(I);
        if (!have_l) P.init_left(I, flav);
    }
}

} // namespace pisab
