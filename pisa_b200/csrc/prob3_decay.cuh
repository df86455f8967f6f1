// prob3_decay.cuh -- the neutrino-decay branch of the propagation (decay_flag == 1).
//
// What the reference computes (numba_osc_kernels.py:445-451, 571-603, 656-685): with decay the Hamiltonian of a layer is
//     H = (H_vac + H_decay) / 2E + H_mat ,   H_decay = U mat_decay U^dagger   (prob3.py:561-563: diag(0, 0, -i alpha3)),
// which is no longer Hermitian; its eigenvalues come from numpy.linalg.eigvals and go through the same Lagrange sum
// (get_product / get_transition_matrix_massbasis) with complex mass differences, so the amplitudes are damped.
//
// B200 formulation: the general-matrix sibling of prob3_device.cuh, in the flavour basis as there --
//   * M = h0 + rho * vm + hd / E with the Hermitian part from the standard tables and the 3x3 complex `hd` of
//     DecayTable; antineutrinos reuse the neutrino code exactly like the standard path (common.cuh): the probabilities
//     of conj(X) equal those of -X, and -X = -hv/E + rho vm + lr + U (-conj(mat_decay)) U^dagger / 2E;
//   * eigenvalues: the trace is removed (A = M - tr/3), the depressed cubic x^3 + p x + q of A is solved in complex
//     arithmetic (Cardano with the cancellation-free choice of the cube root; optional Newton polish, not needed);
//   * exp(-i M t) = a0 + a1 A + a2 A^2 (Cayley-Hamilton form of the same Lagrange sum) times exp(Im(tr/3) t); only a
//     unit-modulus phase is dropped.
// Decay is a per-analysis option, not the fit-loop default: the path shares the fast primitives, the Earth walk and the
// shared-memory state of the Hermitian one but is ~2.8x its FP64 work (27 complex products for A^2, three complex
// exponentials, complex cube / square roots).  Everything is FP64 whatever the storage type.
#pragma once
#include "prob3_device.cuh"

namespace pisab {

// (DecayTable: common.cuh; built by build_decay_table, tables.cu)
// (inlined: as separate functions the 3x3 complex arguments travel through local memory -- 19.3 instead of 15.6 ms per
// 4.8e7 events, profiles/r02_decay.txt)
#define PISAB_DECAY_FN __device__ __forceinline__

__device__ __forceinline__ Cplx cadd(Cplx a, Cplx b) { return Cplx{a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ Cplx csub(Cplx a, Cplx b) { return Cplx{a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ Cplx cscale(double s, Cplx a) { return Cplx{s * a.re, s * a.im}; }
__device__ __forceinline__ double cnorm2(Cplx a) { return fma(a.re, a.re, a.im * a.im); }
// (1 / |b|^2 from the MUFU-seeded reciprocal of prob3_device.cuh: ~1 ulp, no IEEE slow path)
__device__ __forceinline__ Cplx cdiv(Cplx a, Cplx b) {
    const double inv = rcp_fast(cnorm2(b));
    return Cplx{fma(a.re, b.re, a.im * b.im) * inv, fma(a.im, b.re, -a.re * b.im) * inv};
}
// principal square root; |z| from the fast square root (z == 0 gives 0)
__device__ __forceinline__ Cplx csqrt_principal(Cplx z) {
    const double m = sqrt_fast(cnorm2(z));
    const double a2 = 0.5 * (m + fabs(z.re));
    if (!(a2 > 0.0)) return Cplx{0.0, 0.0};
    const double ia = rsqrt_fast(a2);
    const double a = a2 * ia, b = 0.5 * z.im * ia;
    return z.re >= 0.0 ? Cplx{a, b} : Cplx{fabs(b), z.im < 0.0 ? -a : a};
}
// a cube root of z (the principal one): modulus by cbrt, direction by unit_cube_root (float seed + one third-order
// correction, ~1e-18; it wants Im >= 0, the lower half plane goes through the conjugate)
__device__ __forceinline__ Cplx ccbrt_principal(Cplx z) {
    const double m2 = cnorm2(z);
    if (!(m2 > 0.0)) return Cplx{0.0, 0.0};
    const double im = rsqrt_fast(m2);
    const double r = cbrt(m2 * im);
    double c, s;
    unit_cube_root(z.re * im, fabs(z.im) * im, &c, &s);
    return Cplx{r * c, z.im < 0.0 ? -r * s : r * s};
}

// Roots of x^3 + p x + q (complex p, q).
PISAB_DECAY_FN void depressed_cubic_roots(Cplx p, Cplx q, Cplx x[3]) {
    const Cplx hq = cscale(0.5, q), tp = cscale(1.0 / 3.0, p);
    const Cplx disc = cadd(cmul(hq, hq), cmul(cmul(tp, tp), tp));
    const Cplx s = csqrt_principal(disc);
    const Cplx c1 = csub(s, hq), c2 = Cplx{-s.re - hq.re, -s.im - hq.im};
    const Cplx u = ccbrt_principal(cnorm2(c1) >= cnorm2(c2) ? c1 : c2);
    if (cnorm2(u) == 0.0) { // p == q == 0: A == 0
        x[0] = x[1] = x[2] = Cplx{0.0, 0.0};
        return;
    }
    const Cplx v = cdiv(Cplx{-tp.re, -tp.im}, u);
    const Cplx w{-0.5, 0.86602540378443864676}, wc{-0.5, -0.86602540378443864676};
    x[0] = cadd(u, v);
    x[1] = cadd(cmul(u, w), cmul(v, wc));
    x[2] = cadd(cmul(u, wc), cmul(v, w));
#ifndef PISAB_DECAY_NEWTON
#define PISAB_DECAY_NEWTON 0 // Cardano alone matches the oracle to 2e-13, with one or two steps likewise (tests/hostemu)
#endif
#pragma unroll 1
    for (int it = 0; it < PISAB_DECAY_NEWTON; ++it)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const Cplx x2 = cmul(x[k], x[k]);
            const Cplx f = cadd(cmul(cadd(x2, p), x[k]), q);        // (x^2 + p) x + q
            const Cplx d = Cplx{fma(3.0, x2.re, p.re), fma(3.0, x2.im, p.im)};
            if (cnorm2(d) > 0.0) x[k] = csub(x[k], cdiv(f, d));
        }
}

// T = exp(-i M t) up to a unit-modulus phase, M a general complex 3x3 (eV^2/GeV), t = 2 * 2.534 * distance[km].
PISAB_DECAY_FN void transition_matrix_general(const Cplx M[3][3], double t, Mat3 T) {
    const Cplx tr{(M[0][0].re + M[1][1].re + M[2][2].re) * (1.0 / 3.0), (M[0][0].im + M[1][1].im + M[2][2].im) * (1.0 / 3.0)};
    Cplx A[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) A[i][j] = i == j ? csub(M[i][j], tr) : M[i][j];
    // det(x - A) = x^3 + p x + q:  p = sum of the principal 2x2 minors, q = -det A
    const Cplx m01 = csub(cmul(A[0][0], A[1][1]), cmul(A[0][1], A[1][0]));
    const Cplx m02 = csub(cmul(A[0][0], A[2][2]), cmul(A[0][2], A[2][0]));
    const Cplx m12 = csub(cmul(A[1][1], A[2][2]), cmul(A[1][2], A[2][1]));
    const Cplx p = cadd(cadd(m01, m02), m12);
    const Cplx k0 = m12;                                                      // cofactors of row 0
    const Cplx k1 = csub(cmul(A[1][0], A[2][2]), cmul(A[1][2], A[2][0]));
    const Cplx k2 = csub(cmul(A[1][0], A[2][1]), cmul(A[1][1], A[2][0]));
    const Cplx det = cadd(csub(cmul(A[0][0], k0), cmul(A[0][1], k1)), cmul(A[0][2], k2));
    Cplx x[3];
    depressed_cubic_roots(p, Cplx{-det.re, -det.im}, x);
    // Lagrange weights w_k = exp(-i (x_k + i Im tr) t) / prod_{j != k} (x_k - x_j)
    Cplx a0{0.0, 0.0}, a1{0.0, 0.0}, a2{0.0, 0.0};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const Cplx xj = x[(k + 1) % 3], xl = x[(k + 2) % 3];
        double sn, cs;
        sincos_small(x[k].re * t, &sn, &cs);   // |phase| < 1e5 rad, as in the Hermitian path
        const double mag = exp((x[k].im + tr.im) * t);
        const Cplx e{mag * cs, -mag * sn};
        const Cplx w = cdiv(e, cmul(csub(x[k], xj), csub(x[k], xl)));
        a2 = cadd(a2, w);
        a1 = csub(a1, cmul(w, cadd(xj, xl)));
        a0 = cadd(a0, cmul(w, cmul(xj, xl)));
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            Cplx sq = cmul(A[i][0], A[0][j]);
            sq = cfma(A[i][1], A[1][j], sq);
            sq = cfma(A[i][2], A[2][j], sq);
            Cplx v = cfma(a2, sq, cmul(a1, A[i][j]));
            if (i == j) v = cadd(v, a0);
            T[i][j] = v;
        }
}

// Per-event part of the decay Hamiltonian for propagate_earth (prob3_walk.cuh): M0 = h0 + hd / E; a layer adds rho * vm.
struct H0Decay {
    Cplx m0[3][3];
    __device__ __forceinline__ void init(const Herm3 &h, const DecayTable &d, int nubar, double inv_e) {
        const double(*g)[3][2] = d.hd[nubar > 0 ? 0 : 1];
        const Cplx herm[3][3] = {{{h.d0, 0.0}, {h.r01, h.i01}, {h.r02, h.i02}},
                                 {{h.r01, -h.i01}, {h.d1, 0.0}, {h.r12, h.i12}},
                                 {{h.r02, -h.i02}, {h.r12, -h.i12}, {h.d2, 0.0}}};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) m0[i][j] = Cplx{fma(inv_e, g[i][j][0], herm[i][j].re), fma(inv_e, g[i][j][1], herm[i][j].im)};
    }
    __device__ __forceinline__ void layer(double rho, const Herm3 &vm, double t, Mat3 T) const {
        Cplx M[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) M[i][j] = m0[i][j];
        M[0][0].re = fma(rho, vm.d0, M[0][0].re);
        M[1][1].re = fma(rho, vm.d1, M[1][1].re);
        M[2][2].re = fma(rho, vm.d2, M[2][2].re);
        M[0][1] = Cplx{fma(rho, vm.r01, M[0][1].re), fma(rho, vm.i01, M[0][1].im)};
        M[1][0] = Cplx{fma(rho, vm.r01, M[1][0].re), fma(-rho, vm.i01, M[1][0].im)};
        M[0][2] = Cplx{fma(rho, vm.r02, M[0][2].re), fma(rho, vm.i02, M[0][2].im)};
        M[2][0] = Cplx{fma(rho, vm.r02, M[2][0].re), fma(-rho, vm.i02, M[2][0].im)};
        M[1][2] = Cplx{fma(rho, vm.r12, M[1][2].re), fma(rho, vm.i12, M[1][2].im)};
        M[2][1] = Cplx{fma(rho, vm.r12, M[2][1].re), fma(-rho, vm.i12, M[2][1].im)};
        transition_matrix_general(M, t, T);
    }
};
template <>
struct is_general_h0<H0Decay> {
    static constexpr bool value = true;
};

// The same provider with M0 in a per-thread column of shared memory ([18][pitch] doubles, conflict free): frees 36
// registers across the Earth walk of the template kernel.
struct H0DecaySmem {
    double *col; // &s[0][threadIdx.x]
    int pitch;
    static constexpr int kDoubles = 18;
    __device__ __forceinline__ void init(const Herm3 &h, const DecayTable &d, int nubar, double inv_e) {
        H0Decay r;
        r.init(h, d, nubar, inv_e);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                col[((i * 3 + j) * 2) * pitch] = r.m0[i][j].re;
                col[((i * 3 + j) * 2 + 1) * pitch] = r.m0[i][j].im;
            }
    }
    __device__ __forceinline__ void layer(double rho, const Herm3 &vm, double t, Mat3 T) const {
        H0Decay r;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) r.m0[i][j] = Cplx{col[((i * 3 + j) * 2) * pitch], col[((i * 3 + j) * 2 + 1) * pitch]};
        r.layer(rho, vm, t, T);
    }
};
template <>
struct is_general_h0<H0DecaySmem> {
    static constexpr bool value = true;
};

} // namespace pisab
