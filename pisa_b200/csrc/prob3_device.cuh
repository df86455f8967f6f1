// prob3_device.cuh -- per-event three-flavour propagation through constant-density layers.
//
// What the reference computes (numba_osc_kernels.py:121-345,348-531,687-872): for every layer
// with distance > 0 the transition matrix T = exp(-i H t) (H the 3x3 Hermitian Hamiltonian in
// matter, t = 2 * 2.534 * distance), obtained from the eigenvalues of H (closed-form roots of
// the characteristic cubic, get_dms :687-831) and Lagrange's / Sylvester's formula
// (get_product :834-872, get_transition_matrix_massbasis :481-531); the ordered product of the
// layer matrices, rotated to the flavour basis, gives P(i->j) = |A_ji|^2.
//
// B200 formulation (same mathematics, far fewer FP64 operations, no local-memory arrays):
//   * work in the FLAVOUR basis throughout: H = hv/E + rho*vm + lr with per-launch constants,
//     so the two mass-basis rotations per layer (:466-467) and the final rotation (:327-328)
//     disappear;
//   * eigenvalues: the reference's own cubic coefficients c2, c1, c0 and its cancellation-safe
//     discriminant (:716-782), atan2 + ONE sincos (cos(theta +- 2pi/3) by rotation);
//   * exp(-iHt) = a0*1 + a1*H + a2*H^2 (Cayley-Hamilton form of the same Lagrange sum): the
//     27 complex divisions and the 3x3x3 product tensor collapse into one reciprocal, H^2
//     (Hermitian: 9 reals) and three complex coefficients; a global phase is dropped;
//   * Earth symmetry: shell j is crossed on the way in and on the way out with the same
//     (rho, length); T_j is built once and applied to both sides of the running product
//     (what the reference's 1e-5 layer cache achieves with a 17 KiB local array, :224-249);
//   * when only prob_e / prob_mu of one final flavour are needed (fill_probs), a row vector
//     and two column vectors are propagated instead of 3x3 matrices.
#pragma once
#include "common.cuh"

namespace pisab {

struct Cplx {
    double re, im;
};

__device__ __forceinline__ Cplx cmul(Cplx a, Cplx b) {
    return Cplx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
__device__ __forceinline__ Cplx cfma(Cplx a, Cplx b, Cplx c) { // a*b + c
    Cplx r;
    r.re = fma(a.re, b.re, fma(-a.im, b.im, c.re));
    r.im = fma(a.re, b.im, fma(a.im, b.re, c.im));
    return r;
}

typedef Cplx Mat3[3][3];

// Per-event Hamiltonian providers whose layers are general (non-Hermitian) matrices specialise this (prob3_decay.cuh);
// propagate_earth then asks the provider itself for the layer's transition matrix.
template <typename H0>
struct is_general_h0 {
    static constexpr bool value = false;
};

// ---- numeric building blocks ------------------------------------------------------------------
// Non-trivial double constants live in __constant__ memory: a DFMA can take a constant-bank
// operand directly, whereas a 64-bit literal costs two extra MOV/UMOV issue slots per use (ncu on
// the first version of this kernel: 30 % of all issued instructions were such moves).
__constant__ double kTab[24] = {
    /* 0 */ 0.63661977236758134308,   // 2/pi
    /* 1 */ 1.5707963267948966,       // pi/2 hi
    /* 2 */ 6.123233995736766e-17,    // pi/2 lo
    /* 3.. 8: fdlibm __kernel_sin S6..S1 */
    1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,
    -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01,
    /* 9..14: fdlibm __kernel_cos C6..C1 */
    -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,
    2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02,
    /* 15 */ 0.86602540378443864676,  // sin(pi/3)
    /* 16 */ 0.33333333333333333333,  // 1/3
    /* 17 */ 0.66666666666666666667,  // 2/3
    /* 18 */ 5.068,                   // 2 * 2.534  (numba_osc_kernels.py:524)
    /* 19 */ 0.375,
    /* 20 */ 1e-200, // lower clamp of p and of the discriminant
    /* 21 */ 6755399441055744.0, // 1.5 * 2^52: adding it rounds to an integer that sits in the low word
    0, 0};

// 1/x, 1/sqrt(x) and sqrt(x) from the hardware approximations (MUFU.RCP64H / RSQ64H, ~2^-22) refined
// by ONE third-order step (error -> ~2^-64 before rounding, result ~1 ulp): a dependent chain of 3-4
// FP64 instructions instead of the 5-7 of two Newton steps -- these sit on the critical path of every
// layer (the kernel is latency-bound on its FP64 dependency chains, see profiles/).  No IEEE fix-up
// path (the library versions cost ~20 instructions plus a slow-path call).
__device__ __forceinline__ double rcp_seed(double x) {
    double r;
#ifdef PISAB_HOST_EMU
    r = pisab_emu_hi32(1.0 / x); // the hardware result carries only the upper 32 bits
#else
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#endif
    return r;
}
__device__ __forceinline__ double rsqrt_seed(double x) {
    double r;
#ifdef PISAB_HOST_EMU
    r = pisab_emu_hi32(1.0 / sqrt(x));
#else
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#endif
    return r;
}
__device__ __forceinline__ double rcp_fast(double x) {
    const double r = rcp_seed(x);
    const double e = fma(-x, r, 1.0);      // 1/x = r / (1 - e) = r (1 + e + e^2 + O(e^3))
    return fma(r * e, 1.0 + e, r);
}
// with e = 1 - x r^2:  x^(-1/2) = r (1 - e)^(-1/2) = r (1 + e/2 + 3 e^2/8 + O(e^3))
__device__ __forceinline__ double rsqrt_fast(double x) {
    const double r = rsqrt_seed(x);
    const double t = x * r;
    const double e = fma(-t, r, 1.0);
    return fma(r * e, fma(e, kTab[19], 0.5), r);
}
// sqrt(x) = x r (1 + e/2 + 3 e^2/8); x >= 0 (0 allowed)
__device__ __forceinline__ double sqrt_fast(double x) {
    const double xs = fmax(x, 1e-290);
    const double r = rsqrt_seed(xs);
    const double t = xs * r;
    const double e = fma(-t, r, 1.0);
    const double s = fma(t * e, fma(e, kTab[19], 0.5), t);
    return x > 0.0 ? s : 0.0;
}

// sqrt(max(x, tiny)) without the final select (callers for which sqrt(tiny) ~ 1e-100 is as good as 0)
__device__ __forceinline__ double sqrt_pos(double x) {
    const double xs = x > kTab[20] ? x : kTab[20];
    const double r = rsqrt_seed(xs);
    const double t = xs * r;
    const double e = fma(-t, r, 1.0);
    return fma(t * e, fma(e, kTab[19], 0.5), t);
}

// sin/cos for |x| < ~1e5 rad (phases here are < 1e3): two-term Cody-Waite reduction with FMA and
// the fdlibm kernel polynomials on [-pi/4, pi/4].  No Payne-Hanek slow path (keeps the code small
// enough for the instruction cache); max error ~1 ulp, same as the CUDA fast path.
__device__ __forceinline__ void sincos_small(double x, double *sn, double *cs) {
#ifdef PISAB_SINCOS_RINT
    const double kd = rint(x * kTab[0]);
    const int k = (int)kd;
#else
    // round-to-nearest by the 1.5 * 2^52 shift (|x * 2/pi| < 2^31 here): the integer is the low word of the
    // shifted sum, so no FRND / F2I conversions sit on the dependent chain
    const double shifted = fma(x, kTab[0], kTab[21]);
    const int k = __double2loint(shifted);
    const double kd = shifted - kTab[21];
#endif
    double r = fma(-kd, kTab[1], x);
    r = fma(-kd, kTab[2], r);
    const double z = r * r;
    // Horner: one constant-bank operand per FMA (an Estrin split needs two constants in one FMA, i.e. an
    // extra LDC each, and the kernel is issue-bound, not latency-bound: measured 1.5 % slower)
    double ps = fma(z, kTab[3], kTab[4]);
    ps = fma(z, ps, kTab[5]);
    ps = fma(z, ps, kTab[6]);
    ps = fma(z, ps, kTab[7]);
    ps = fma(z, ps, kTab[8]);
    const double s = fma(r * z, ps, r);
    double pc = fma(z, kTab[9], kTab[10]);
    pc = fma(z, pc, kTab[11]);
    pc = fma(z, pc, kTab[12]);
    pc = fma(z, pc, kTab[13]);
    pc = fma(z, pc, kTab[14]);
    const double c = fma(z * z, pc, fma(z, -0.5, 1.0));
    const double ss = (k & 1) ? c : s;
    const double cc = (k & 1) ? s : c;
#ifdef PISAB_SIGN_SELECT
    *sn = (k & 2) ? -ss : ss;
    *cs = ((k + 1) & 2) ? -cc : cc;
#else
    // quadrant signs as sign-bit XORs (bit 1 of k resp. k + 1 moved to bit 31 of the high word): three integer
    // instructions per value instead of a negate and two selects
    *sn = __hiloint2double(__double2hiint(ss) ^ ((k << 30) & 0x80000000), __double2loint(ss));
    *cs = __hiloint2double(__double2hiint(cc) ^ (((k + 1) << 30) & 0x80000000), __double2loint(cc));
#endif
}

// (cos t, sin t) with 3t = atan2(zi, zr), zi >= 0, i.e. the principal cube root of the unit complex
// number z/|z| (replaces atan2 + sincos of numba_osc_kernels.py:794-813).  A float-precision seed
// (polynomial atan + sincos, ~5e-7) is normalised and refined by one third-order angle correction
//     d = Im(z conj(w^3)) / 3 ,  w <- w (1 - d^2/2 + i d)
// which leaves an error ~1.5 d^3 ~ 1e-18.  The conditioning of the eigenvalues sits entirely in
// (zr, zi), not in this step.
__device__ __forceinline__ void unit_cube_root_seed(double zr, double zi, float *c_seed, float *s_seed) {
    // (zr, zi) is already of unit modulus (up to rounding): the caller scales by p^(-3/2), because
    // q^2 + (p^3 - q^2) = p^3 -- no second rsqrt here.  A modulus error delta only enters the
    // correction step as delta * d ~ 1e-23.
    // float seed of theta = atan2(zi, zr) / 3, zi >= 0, good to ~5e-7 rad: octant reduction, a degree-6
    // minimax polynomial for atan(t)/t on [0, 1] (5.5e-7), then sin / cos polynomials valid on [0, pi/3]
    // (no range reduction needed) -- 27 instructions instead of the 45 of atan2f + __sincosf.
    const float x = (float)zr, y = (float)zi;
    const float ax = fabsf(x);
    const float mx = fmaxf(ax, y), mn = fminf(ax, y);
#ifdef PISAB_HOST_EMU
    const float t = mn / mx;
#else
    const float t = __fdividef(mn, mx);
#endif
    const float u = t * t;
    float a = fmaf(u, 0.00782548263669014f, -0.03689862787723541f);
    a = fmaf(u, a, 0.08374155312776566f);
    a = fmaf(u, a, -0.13480405509471893f);
    a = fmaf(u, a, 0.19879871606826782f);
    a = fmaf(u, a, -0.3332637548446655f);
    a = fmaf(u, a, 0.9999993443489075f) * t;
    a = y > ax ? 1.57079632679489662f - a : a;
    a = x < 0.0f ? 3.14159265358979324f - a : a;
    const float th = a * (1.0f / 3.0f), v = th * th;
    *s_seed = th * fmaf(v, fmaf(v, fmaf(v, -0.00019222621631342918f, 0.008328950963914394f), -0.1666656732559204f),
                        0.9999999403953552f);
    *c_seed = fmaf(v, fmaf(v, fmaf(v, -0.0013333901297301054f, 0.041627395898103714f), -0.4999910891056061f),
                   0.9999997019767761f);
}
__device__ __forceinline__ void unit_cube_root(double zr, double zi, double *c_out, double *s_out) {
    float sf, cf;
    unit_cube_root_seed(zr, zi, &cf, &sf);
    const double c = (double)cf, s = (double)sf;
    // u = (c, s) has |u|^2 = 1 + m, m ~ 1e-7.  The normalisation (1+m)^(-1/2) = rho and the angle
    // correction are evaluated as two independent chains (depth 9 instead of 15):
    //   d = Im(z conj(w^3)) / 3 = Im(z conj(u^3)) (1 - 3m/2) / 3 ,  w' = rho u (1 - d^2/2 + i d)
    const double m = fma(c, c, fma(s, s, -1.0));
    const double rho = fma(m, fma(m, kTab[19], -0.5), 1.0);
    const double cr = c * rho, sr = s * rho;
    const double third = fma(m, -0.5, kTab[16]); // (1 - 1.5 m) / 3
    const double u2r = fma(c, c, -s * s), u2i = 2.0 * c * s;
    const double u3r = fma(u2r, c, -u2i * s), u3i = fma(u2r, s, u2i * c);
    const double d = fma(zi, u3r, -zr * u3i) * third;
    const double k = fma(-0.5 * d, d, 1.0);
    *c_out = fma(cr, k, -sr * d);
    *s_out = fma(sr, k, cr * d);
}
// FP32-mode variant: first-order angle correction and first-order normalisation, w' = (1 - m/2) u (1 + i d).  The
// neglected terms are d^2 / 2 ~ 1e-13 and 3 m^2 / 8 ~ 4e-15 -- far below the 1e-9 the eigenvalue gaps need there.
__device__ __forceinline__ void unit_cube_root_mp(double zr, double zi, double *c_out, double *s_out) {
    float sf, cf;
    unit_cube_root_seed(zr, zi, &cf, &sf);
    const double c = (double)cf, s = (double)sf;
    const double m = fma(c, c, fma(s, s, -1.0));
    const double rho = fma(m, -0.5, 1.0);
    const double third = fma(m, -0.5, kTab[16]);
    const double u2r = fma(c, c, -s * s), u2i = 2.0 * c * s;
    const double u3r = fma(u2r, c, -u2i * s), u3i = fma(u2r, s, u2i * c);
    const double d = fma(zi, u3r, -zr * u3i) * third;
    *c_out = fma(-s, d, c) * rho;
    *s_out = fma(c, d, s) * rho;
}

// H = hv * inv_e + lr   (per event), then + rho * vm per layer
__device__ __forceinline__ Herm3 herm_axpy(double a, const Herm3 &x, const Herm3 &y) {
    Herm3 r;
    r.d0 = fma(a, x.d0, y.d0);
    r.d1 = fma(a, x.d1, y.d1);
    r.d2 = fma(a, x.d2, y.d2);
    r.r01 = fma(a, x.r01, y.r01);
    r.i01 = fma(a, x.i01, y.i01);
    r.r02 = fma(a, x.r02, y.r02);
    r.i02 = fma(a, x.i02, y.i02);
    r.r12 = fma(a, x.r12, y.r12);
    r.i12 = fma(a, x.i12, y.i12);
    return r;
}

// Characteristic polynomial x^3 + c2 x^2 + c1 x + c0 of a Hermitian 3x3 (numba_osc_kernels.py:716-752)
__device__ __forceinline__ void char_poly(const Herm3 &h, double &c2, double &c1, double &c0) {
    const double n01 = fma(h.r01, h.r01, h.i01 * h.i01);
    const double n02 = fma(h.r02, h.r02, h.i02 * h.i02);
    const double n12 = fma(h.r12, h.r12, h.i12 * h.i12);
    const double ur = fma(h.r01, h.r12, -h.i01 * h.i12); // h01*h12
    const double ui = fma(h.r01, h.i12, h.i01 * h.r12);
    const double rpa = fma(ur, h.r02, ui * h.i02);       // Re(h01 h12 h20)
    c2 = -(h.d0 + h.d1 + h.d2);
    c1 = fma(h.d0, h.d1 + h.d2, h.d1 * h.d2) - n01 - n12 - n02;
    c0 = fma(h.d0, n12, fma(h.d1, n02, h.d2 * n01)) - 2.0 * rpa - h.d0 * h.d1 * h.d2;
}

// Roots of x^3 + c2 x^2 + c1 x + c0 (three real roots: the matrix is Hermitian) and the Lagrange
// denominators, numba_osc_kernels.py:766-814 / :870-872.  They depend on the layer's density but not on
// its length, so the two pieces of the detector shell share one solve.
struct Eigen {
    double l0, l1, l2;    // ascending
    double id0, id1, id2; // 1 / prod_{j != k} (l_k - l_j)
};
__device__ __forceinline__ Eigen eigen_solve(double c2, double c1, double c0) {
    // p, q and the cancellation-safe p^3 - q^2.  Both are clamped from below with a plain compare-select
    // (fmax's NaN handling costs 5 instructions); p == 0 would mean H proportional to 1 (three equal roots),
    // which no oscillation Hamiltonian with a non-zero mass splitting or matter term produces.
    double p = fma(c2, c2, -3.0 * c1);
    p = p > kTab[20] ? p : kTab[20];
    const double q = fma(4.5 * c1, c2, fma(-13.5, c0, -c2 * c2 * c2));
    const double disc = 27.0 * fma(0.25 * c1 * c1, p - c1, c0 * fma(6.75, c0, q));
    // sqrt(p) and p^(-3/2) from one rsqrt
    const double rs = rsqrt_fast(p);
    const double b = kTab[17] * (p * rs);
    const double inv = rs * rs * rs;
    const double base = -c2 * kTab[16];
    double st, ct;
    // theta = atan2(sqrt(disc), q) / 3 in [0, pi/3]
    unit_cube_root(q * inv, sqrt_pos(disc) * inv, &ct, &st);
    const double kh = 0.5, ks = kTab[15]; // cos, sin of pi/3
    Eigen e;
    // theta+2pi/3 -> smallest root, theta-2pi/3 -> middle, theta -> largest (:795-797)
    e.l0 = fma(b, -kh * ct - ks * st, base);
    e.l1 = fma(b, -kh * ct + ks * st, base);
    e.l2 = fma(b, ct, base);
    const double g01 = e.l0 - e.l1, g02 = e.l0 - e.l2, g12 = e.l1 - e.l2;
    const double inv_g = rcp_fast(g01 * g02 * g12);
    e.id0 = g12 * inv_g;  // 1/((l0-l1)(l0-l2))
    e.id1 = -g02 * inv_g; // 1/((l1-l0)(l1-l2))
    e.id2 = g01 * inv_g;  // 1/((l2-l0)(l2-l1))
    return e;
}

// H^2 of a Hermitian 3x3, packed like H (diagonal real, upper triangle complex)
__device__ __forceinline__ Herm3 herm_square(const Herm3 &h) {
    const double n01 = fma(h.r01, h.r01, h.i01 * h.i01);
    const double n02 = fma(h.r02, h.r02, h.i02 * h.i02);
    const double n12 = fma(h.r12, h.r12, h.i12 * h.i12);
    const double t01 = h.d0 + h.d1, t02 = h.d0 + h.d2, t12 = h.d1 + h.d2;
    Herm3 s;
    s.d0 = fma(h.d0, h.d0, n01 + n02);
    s.d1 = fma(h.d1, h.d1, n01 + n12);
    s.d2 = fma(h.d2, h.d2, n02 + n12);
    s.r01 = fma(h.r01, t01, fma(h.r02, h.r12, h.i02 * h.i12));
    s.i01 = fma(h.i01, t01, fma(h.i02, h.r12, -h.r02 * h.i12));
    s.r02 = fma(h.r02, t02, fma(h.r01, h.r12, -h.i01 * h.i12));
    s.i02 = fma(h.i02, t02, fma(h.r01, h.i12, h.i01 * h.r12));
    s.r12 = fma(h.r12, t12, fma(h.r01, h.r02, h.i01 * h.i02));
    s.i12 = fma(h.i12, t12, fma(h.r01, h.i02, -h.i01 * h.r02));
    return s;
}

// T = exp(-i H t) up to a global phase, H = h + x e00 e00^T, from the roots of H, h and h^2.
// exp(-iHt) = a0 + a1 H + a2 H^2 (Cayley-Hamilton form of the Lagrange sum :481-531,834-872) and
//     H^2 = h^2 + x (e00 h + h e00) + x^2 e00 ,
// so only row / column 0 see the shift: with b1 = a1 + a2 x they use (b1, a2) where the rest uses (a1, a2).
// SHIFT = false (x == 0): h, hsq are the layer's own H and H^2.  `t` = 2 * 2.534 * distance[km].
template <bool SHIFT>
__device__ __forceinline__ void assemble_transition(const Herm3 &h, const Herm3 &hsq, double x, const Eigen &e,
                                                    double t, Mat3 T) {
    // ---- Lagrange weights with the global phase exp(-i l2 t) dropped
    double s0, k0, s1, k1;
    sincos_small((e.l2 - e.l0) * t, &s0, &k0); // exp(-i (l0-l2) t)
    sincos_small((e.l2 - e.l1) * t, &s1, &k1); // exp(-i (l1-l2) t)
    const Cplx w0{k0 * e.id0, s0 * e.id0};
    const Cplx w1{k1 * e.id1, s1 * e.id1};
    const double w2 = e.id2;
    const double m12 = e.l1 + e.l2, m02 = e.l0 + e.l2, m01 = e.l0 + e.l1;
    const double p12 = e.l1 * e.l2, p02 = e.l0 * e.l2, p01 = e.l0 * e.l1;
    const Cplx a2{w0.re + w1.re + w2, w0.im + w1.im};
    const Cplx a1{-fma(w0.re, m12, fma(w1.re, m02, w2 * m01)), -fma(w0.im, m12, w1.im * m02)};
    const Cplx a0{fma(w0.re, p12, fma(w1.re, p02, w2 * p01)), fma(w0.im, p12, w1.im * p02)};
    const Cplx b1 = SHIFT ? Cplx{fma(a2.re, x, a1.re), fma(a2.im, x, a1.im)} : a1;
    const double d0 = SHIFT ? h.d0 + x : h.d0;
    const double q0 = SHIFT ? fma(x, fma(2.0, h.d0, x), hsq.d0) : hsq.d0;

    // ---- assemble T
    T[0][0] = Cplx{fma(a2.re, q0, fma(a1.re, d0, a0.re)), fma(a2.im, q0, fma(a1.im, d0, a0.im))};
    T[1][1] = Cplx{fma(a2.re, hsq.d1, fma(a1.re, h.d1, a0.re)), fma(a2.im, hsq.d1, fma(a1.im, h.d1, a0.im))};
    T[2][2] = Cplx{fma(a2.re, hsq.d2, fma(a1.re, h.d2, a0.re)), fma(a2.im, hsq.d2, fma(a1.im, h.d2, a0.im))};
#define PISAB_OFFDIAG(I, J, C1, RE, IM, SR, SI)                               \
    {                                                                         \
        const double xr = fma(a2.re, SR, C1.re * RE), xi = fma(a2.im, SR, C1.im * RE); \
        const double yr = fma(a2.re, SI, C1.re * IM), yi = fma(a2.im, SI, C1.im * IM); \
        T[I][J] = Cplx{xr - yi, xi + yr};                                     \
        T[J][I] = Cplx{xr + yi, xi - yr};                                     \
    }
    PISAB_OFFDIAG(0, 1, b1, h.r01, h.i01, hsq.r01, hsq.i01)
    PISAB_OFFDIAG(0, 2, b1, h.r02, h.i02, hsq.r02, hsq.i02)
    PISAB_OFFDIAG(1, 2, a1, h.r12, h.i12, hsq.r12, hsq.i12)
#undef PISAB_OFFDIAG
}

// T = exp(-i h t) up to a global phase.  `h` in eV^2/GeV, t = 2 * 2.534 * distance[km].
__device__ __forceinline__ void transition_matrix(const Herm3 &h, double t, Mat3 T) {
    double c2, c1, c0;
    char_poly(h, c2, c1, c0);
    const Eigen e = eigen_solve(c2, c1, c0);
    assemble_transition<false>(h, herm_square(h), 0.0, e, t, T);
}

// ---- small dense helpers on NR x 3 / 3 x NC blocks ----------------------------------------
// L (NR x 3) <- L . T
template <int NR>
__device__ __forceinline__ void left_times(Cplx (*L)[3], const Mat3 T) {
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        Cplx o[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            Cplx acc = cmul(L[r][0], T[0][c]);
            acc = cfma(L[r][1], T[1][c], acc);
            acc = cfma(L[r][2], T[2][c], acc);
            o[c] = acc;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) L[r][c] = o[c];
    }
}
// R (3 x NC, stored as R[c][k] = column c, component k) <- T . R
template <int NC>
__device__ __forceinline__ void times_right(const Mat3 T, Cplx (*R)[3]) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        Cplx o[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            Cplx acc = cmul(T[k][0], R[c][0]);
            acc = cfma(T[k][1], R[c][1], acc);
            acc = cfma(T[k][2], R[c][2], acc);
            o[k] = acc;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) R[c][k] = o[k];
    }
}

// ---- Earth geometry (layers.py:38-169) with the reference's operation order ----------------
// sqrt argument r_det^2*cz^2 - r_det^2 + r_j^2, evaluated without FMA contraction so that the
// distances are bit-identical to numpy's.
__device__ __forceinline__ double shell_root(double rd2, double cz2, double rj2) {
    return __dsqrt_rn(__dadd_rn(__dsub_rn(__dmul_rn(rd2, cz2), rd2), rj2));
}

// Same argument, ~1 ulp square root without the IEEE fix-up path: used by the propagation kernels,
// where a 1e-16 relative change of a segment length moves a phase by < 1e-13 (the stand-alone layers
// kernel keeps the correctly rounded version and stays bit-identical to numpy).
__device__ __forceinline__ double shell_root_fast(double rd2, double cz2, double rj2) {
#ifdef PISAB_EXACT_SQRT
    return shell_root(rd2, cz2, rj2);
#else
    return sqrt_pos(__dadd_rn(__dsub_rn(__dmul_rn(rd2, cz2), rd2), rj2)); // (a crossed shell has a positive argument)
#endif
}

// Propagation state: L is NR x 3 (rows of the product on the detector side), R holds NC
// columns of the product on the production side.
//   MODE_FULL: NR = 3, NC = 3  -> probability[3][3]
//   MODE_ROW : NR = 1, NC = 2  -> prob_e, prob_mu of final flavour `flav`
// Two storage policies with the same interface:
//   Propagator      : registers (18 / 36 doubles live across the eigenvalue solve);
//   PropagatorSmem  : a per-thread column of shared memory, [(v*3+k)*2+re/im][block] doubles, so that
//                     the vectors occupy registers only while they are being multiplied -- the
//                     eigenvalue solve + matrix assembly then fit the 128-register budget without
//                     spills, and the loop carries no register moves for the state.
template <int NR, int NC>
struct Propagator {
    static constexpr bool kF32 = false;
    typedef Cplx cplx;
    Cplx L[NR][3];
    Cplx R[NC][3];

    __device__ __forceinline__ void set_right(int c, int k, Cplx v) { R[c][k] = v; }
    __device__ __forceinline__ void init_right(const Mat3 T) {
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int k = 0; k < 3; ++k) R[c][k] = T[k][c];
    }
    __device__ __forceinline__ void init_left(const Mat3 T, int flav) {
        if (NR == 3) {
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) L[r][c] = T[r][c];
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                Cplx v = T[0][c];
                if (flav == 1) v = T[1][c];
                if (flav == 2) v = T[2][c];
                L[0][c] = v;
            }
        }
    }
    __device__ __forceinline__ void mul_right(const Mat3 T) { times_right<NC>(T, R); }
    __device__ __forceinline__ void mul_left(const Mat3 T) { left_times<NR>(L, T); }
    // amplitude A[r][c] = sum_k L[r][k] R[c][k]; returns |A|^2
    __device__ __forceinline__ double prob(int r, int c) const {
        Cplx acc = cmul(L[r][0], R[c][0]);
        acc = cfma(L[r][1], R[c][1], acc);
        acc = cfma(L[r][2], R[c][2], acc);
        return fma(acc.re, acc.re, acc.im * acc.im);
    }
};

template <int NR, int NC>
struct PropagatorSmem {
    static constexpr bool kF32 = false;
    typedef Cplx cplx;
    double2 *col; // this thread's column: &state[0][threadIdx.x]; one (re, im) pair per 128-bit access
    int pitch;    // entries between consecutive rows (= block size)
    static constexpr int kDoubles = (NR + NC) * 6;

    __device__ __forceinline__ Cplx ld(int v, int k) const {
        const double2 z = col[(v * 3 + k) * pitch];
        return Cplx{z.x, z.y};
    }
    __device__ __forceinline__ void st(int v, int k, Cplx z) { col[(v * 3 + k) * pitch] = make_double2(z.re, z.im); }
    // vectors 0..NC-1 = columns of R, NC..NC+NR-1 = rows of L
    __device__ __forceinline__ void set_right(int c, int k, Cplx v) { st(c, k, v); }
    __device__ __forceinline__ void init_right(const Mat3 T) {
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int k = 0; k < 3; ++k) st(c, k, T[k][c]);
    }
    __device__ __forceinline__ void init_left(const Mat3 T, int flav) {
        if (NR == 3) {
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) st(NC + r, c, T[r][c]);
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                Cplx v = T[0][c];
                if (flav == 1) v = T[1][c];
                if (flav == 2) v = T[2][c];
                st(NC, c, v);
            }
        }
    }
    __device__ __forceinline__ void mul_right(const Mat3 T) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const Cplx r0 = ld(c, 0), r1 = ld(c, 1), r2 = ld(c, 2);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                Cplx acc = cmul(T[k][0], r0);
                acc = cfma(T[k][1], r1, acc);
                acc = cfma(T[k][2], r2, acc);
                st(c, k, acc);
            }
        }
    }
    __device__ __forceinline__ void mul_left(const Mat3 T) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const Cplx l0 = ld(NC + r, 0), l1 = ld(NC + r, 1), l2 = ld(NC + r, 2);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                Cplx acc = cmul(l0, T[0][c]);
                acc = cfma(l1, T[1][c], acc);
                acc = cfma(l2, T[2][c], acc);
                st(NC + r, c, acc);
            }
        }
    }
    __device__ __forceinline__ double prob(int r, int c) const {
        Cplx acc = cmul(ld(NC + r, 0), ld(c, 0));
        acc = cfma(ld(NC + r, 1), ld(c, 1), acc);
        acc = cfma(ld(NC + r, 2), ld(c, 2), acc);
        return fma(acc.re, acc.re, acc.im * acc.im);
    }
};

// Columns 0..NC-1 of the vacuum transition matrix 1 + z2 P2 + z3 P3 (see OscTable), written
// straight into R (R[c][k] = T[k][c]); ts = -/+ t / E for nu / nubar.
template <int NC, typename PROP>
__device__ __forceinline__ void vacuum_columns(const OscTable &o, double ts, PROP &P) {
    double s2, c2, s3, c3;
    sincos_small(o.hdm21 * ts, &s2, &c2);
    sincos_small(o.hdm31 * ts, &s3, &c3);
    const double z2r = c2 - 1.0, z2i = s2, z3r = c3 - 1.0, z3i = s3;
    const Herm3 &A = o.pr2, &B = o.pr3;
#define PISAB_VAC_DIAG(C, DA, DB) \
    P.set_right(C, C, Cplx{fma(z2r, DA, fma(z3r, DB, 1.0)), fma(z2i, DA, z3i * DB)});
#define PISAB_VAC_OFF(I, J, AR, AI, BR, BI)                                         \
    {                                                                               \
        const double xr = fma(z2r, AR, z3r * BR), xi = fma(z2i, AR, z3i * BR);        \
        const double yr = fma(z2r, AI, z3r * BI), yi = fma(z2i, AI, z3i * BI);        \
        if (J < NC) P.set_right(J, I, Cplx{xr - yi, xi + yr}); /* T[I][J] */           \
        if (I < NC) P.set_right(I, J, Cplx{xr + yi, xi - yr}); /* T[J][I] */           \
    }
    PISAB_VAC_DIAG(0, A.d0, B.d0)
    PISAB_VAC_DIAG(1, A.d1, B.d1)
    if (NC > 2) PISAB_VAC_DIAG(2, A.d2, B.d2)
    PISAB_VAC_OFF(0, 1, A.r01, A.i01, B.r01, B.i01)
    PISAB_VAC_OFF(0, 2, A.r02, A.i02, B.r02, B.i02)
    PISAB_VAC_OFF(1, 2, A.r12, A.i12, B.r12, B.i12)
#undef PISAB_VAC_DIAG
#undef PISAB_VAC_OFF
}

// Where the per-event part of the Hamiltonian (h0 = hv/E + lr, 9 doubles) lives between layers:
// in registers, or in a per-thread column of shared memory (frees 18 registers for the
// eigenvalue solve; the 9 LDS per layer are conflict-free with the [9][block] layout).
// For the standard matter potential (vm = diag(a, 0, 0): no NSI) a layer changes only H[0][0], by
// x = rho * a, and the characteristic polynomial is linear in x:
//     c2 = c2_0 - x ,  c1 = c1_0 + x (d1 + d2) ,  c0 = c0_0 - x (d1 d2 - |h12|^2)
// so the provider also keeps the 5 per-event invariants and a layer costs 3 FMAs instead of 17 + 9.
struct H0Reg {
    Herm3 h, hsq;
    double inv[5]; // c2_0, c1_0, c0_0, d1 + d2, d1 d2 - |h12|^2 (only set by set_poly)
    __device__ __forceinline__ Herm3 load() const { return h; }
    __device__ __forceinline__ Herm3 load_sq() const { return hsq; }
    __device__ __forceinline__ void set_poly() {
        char_poly(h, inv[0], inv[1], inv[2]);
        inv[3] = h.d1 + h.d2;
        inv[4] = fma(h.d1, h.d2, -fma(h.r12, h.r12, h.i12 * h.i12));
        hsq = herm_square(h);
    }
    __device__ __forceinline__ void poly(double x, double &c2, double &c1, double &c0) const {
        c2 = inv[0] - x;
        c1 = fma(x, inv[3], inv[1]);
        c0 = fma(-x, inv[4], inv[2]);
    }
};
template <bool STD>
struct H0Smem {
    // Per-thread column in shared memory, stored as double2 rows (one 128-bit access per pair, conflict free)
    // plus one trailing double row:
    //   general : (d0,d1) (d2,r01) (i01,r02) (i02,r12) | i12                                  = 9 doubles
    //   STD     : (d0,d1) (d2,r01) (i01,r02) (i02,r12) (i12,c2_0) (c1_0,c0_0) (d1+d2, m00)
    //             (q.d0,q.d1) (q.d2,q.r01) (q.i01,q.r02) (q.i02,q.r12) | q.i12   (q = h0^2)   = 23 doubles
    double *col; // &s_h0[0][threadIdx.x] of a [kDoubles][pitch] double block
    int pitch;   // block size
    static constexpr int kPairs = STD ? 11 : 4;
    static constexpr int kDoubles = 2 * kPairs + 1;
    __device__ __forceinline__ double2 &pair(int r) const { return reinterpret_cast<double2 *>(col - threadIdx.x)[r * pitch + threadIdx.x]; }
    __device__ __forceinline__ double &single() const { return (col - threadIdx.x)[2 * kPairs * pitch + threadIdx.x]; }

    __device__ __forceinline__ void store(const Herm3 &h) {
        pair(0) = make_double2(h.d0, h.d1);
        pair(1) = make_double2(h.d2, h.r01);
        pair(2) = make_double2(h.i01, h.r02);
        pair(3) = make_double2(h.i02, h.r12);
        if (!STD) single() = h.i12;
    }
    __device__ __forceinline__ Herm3 load() const {
        Herm3 h;
        const double2 a = pair(0), b = pair(1), c = pair(2), e = pair(3);
        h.d0 = a.x; h.d1 = a.y; h.d2 = b.x; h.r01 = b.y; h.i01 = c.x; h.r02 = c.y; h.i02 = e.x; h.r12 = e.y;
        h.i12 = STD ? pair(4).x : single();
        return h;
    }
    __device__ __forceinline__ void set_poly(const Herm3 &h) {
        double c2, c1, c0;
        char_poly(h, c2, c1, c0);
        pair(4) = make_double2(h.i12, c2);
        pair(5) = make_double2(c1, c0);
        pair(6) = make_double2(h.d1 + h.d2, fma(h.d1, h.d2, -fma(h.r12, h.r12, h.i12 * h.i12)));
        const Herm3 s = herm_square(h);
        pair(7) = make_double2(s.d0, s.d1);
        pair(8) = make_double2(s.d2, s.r01);
        pair(9) = make_double2(s.i01, s.r02);
        pair(10) = make_double2(s.i02, s.r12);
        single() = s.i12;
    }
    __device__ __forceinline__ Herm3 load_sq() const {
        Herm3 s;
        const double2 a = pair(7), b = pair(8), c = pair(9), e = pair(10);
        s.d0 = a.x; s.d1 = a.y; s.d2 = b.x; s.r01 = b.y; s.i01 = c.x; s.r02 = c.y; s.i02 = e.x; s.r12 = e.y;
        s.i12 = single();
        return s;
    }
    __device__ __forceinline__ void poly(double x, double &c2, double &c1, double &c0) const {
        const double2 u = pair(5), v = pair(6);
        c2 = pair(4).y - x;
        c1 = fma(x, v.x, u.x);
        c0 = fma(-x, v.y, u.y);
    }
};

} // namespace pisab
