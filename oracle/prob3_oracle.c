/*
 * oracle/prob3_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, op-for-op restatement of the reference's numba CPU path for
 * three-flavour oscillation probabilities through constant-density layers
 * (icecube/pisa, pisa/stages/osc/prob3numba/numba_osc_kernels.py and
 * numba_osc_hostfuncs.py).  It exists to *check* the CUDA path in tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 * Nothing under pisa_b200/ may call it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function
 * below against the reference's own golden pickles
 * (pisa_examples/resources/osc/numba_osc_tests_data, re-packed by
 * tests/golden/make_golden.py) and against outputs of the unmodified reference
 * run in the build container (tests/golden/ref_*.npz).
 *
 * Arithmetic notes (why this is compiled with -ffp-contract=off):
 *  - numba compiles the reference without fastmath, so there is no FMA
 *    contraction and every + - * / sqrt is an IEEE double op; libm supplies
 *    cos / sin / atan2.  Complex multiply is the naive 4-mul form, complex
 *    division by a real-valued complex is a per-component real division
 *    (CPython's _Py_c_quot algorithm with bimag == 0).
 *  - FP64 ONLY.  With PISA_FTYPE=fp32 numba types every Python float literal
 *    (2.534, 1.52588e-4, 0.5, 3.0, ...) as float64 and promotes float32 (op)
 *    float64 -> float64, so the reference's "FP32" path is an accidental mix
 *    (float32 storage, mostly float64 arithmetic in get_dms / the phases) that
 *    differs from its own FP64 result by up to 7e-5.  It is not
 *    restated; FP32-mode parity is defined against THIS FP64 oracle evaluated
 *    on the float32-rounded inputs (tolerance 1e-5 absolute, BASELINE.json),
 *    and the distance to the reference's f4 fixtures is reported beside it.
 *
 * The decay branch (decay_flag == 1, numba_osc_kernels.py:445-451,656-685) calls
 * numpy.linalg.eigvals, i.e. LAPACK zgeev (numpy >= 1.21 bundled OpenBLAS; not
 * under /root/reference).  It is restated here as what zgeev does for the
 * eigenvalues of a general complex matrix: reduction to upper Hessenberg form
 * followed by the single-shift (Wilkinson) QR iteration with deflation
 * (eigvals3 below).  The order of the eigenvalues is not part of the contract:
 * the Lagrange sum of get_transition_matrix_massbasis is symmetric in them.
 * Pinned against the reference's nufit32_std_decay pickles and against
 * outputs of the unmodified reference (tests/golden/ref_decay_f8.npz).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifdef ORACLE_F32
typedef float real_t;
#define R_SQRT sqrtf
#define R_COS cosf
#define R_SIN sinf
#define R_ATAN2 atan2f
#define R_FABS fabsf
#define SUFFIX(name) name##_f32
#else
typedef double real_t;
#define R_SQRT sqrt
#define R_COS cos
#define R_SIN sin
#define R_ATAN2 atan2
#define R_FABS fabs
#define SUFFIX(name) name##_f64
#endif

typedef struct { real_t re, im; } cplx;

#define MAX_LAYERS 120 /* numba_osc_kernels.py:227 */

static inline cplx c_make(real_t re, real_t im) { cplx z; z.re = re; z.im = im; return z; }
static inline cplx c_add(cplx a, cplx b) { return c_make(a.re + b.re, a.im + b.im); }
static inline cplx c_sub(cplx a, cplx b) { return c_make(a.re - b.re, a.im - b.im); }
static inline cplx c_mul(cplx a, cplx b) {
    return c_make(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
static inline cplx c_conj(cplx a) { return c_make(a.re, -a.im); }
/* complex division as numba / CPython do it (Smith's method, _Py_c_quot); with b.im == 0 this is
 * the per-component real division */
static inline cplx c_div(cplx a, cplx b) {
    real_t abs_breal = b.re < 0 ? -b.re : b.re, abs_bimag = b.im < 0 ? -b.im : b.im;
    if (abs_breal >= abs_bimag) {
        if (abs_breal == (real_t)0) return c_make((real_t)NAN, (real_t)NAN);
        real_t ratio = b.im / b.re, denom = b.re + b.im * ratio;
        return c_make((a.re + a.im * ratio) / denom, (a.im - a.re * ratio) / denom);
    } else {
        real_t ratio = b.re / b.im, denom = b.re * ratio + b.im;
        return c_make((a.re * ratio + a.im) / denom, (a.im * ratio - a.re) / denom);
    }
}
/* numba promotes a real factor to (x + 0j) before multiplying */
static inline cplx c_rmul(real_t x, cplx a) { return c_mul(c_make(x, (real_t)0), a); }

/* numba_tools.py:278-289 matrix_dot_matrix: C = A.B, j outer, i inner, n innermost */
static void mat_mul(const cplx A[3][3], const cplx B[3][3], cplx C[3][3]) {
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i) {
            cplx acc = c_make(0, 0);
            for (int n = 0; n < 3; ++n) acc = c_add(acc, c_mul(A[i][n], B[n][j]));
            C[i][j] = acc;
        }
}

static void mat_copy(const cplx A[3][3], cplx B[3][3]) { memcpy(B, A, sizeof(cplx) * 9); }

/* numba_osc_kernels.py:534-569 get_H_vac */
void SUFFIX(oracle_get_H_vac)(const cplx mix[3][3], const cplx mix_ct[3][3],
                              const real_t dm[3][3], cplx H_vac[3][3]) {
    cplx diag[3][3], tmp[3][3];
    memset(diag, 0, sizeof diag);
    diag[1][1] = c_make(dm[1][0], 0);
    diag[2][2] = c_make(dm[2][0], 0);
    mat_mul(diag, mix_ct, tmp);
    mat_mul(mix, tmp, H_vac);
}

/* numba_osc_kernels.py:605-653 get_H_mat */
void SUFFIX(oracle_get_H_mat)(real_t rho, const cplx mat_pot[3][3], int64_t nubar,
                              cplx H_mat[3][3]) {
    const real_t tworttwoGf = (real_t)1.52588e-4;
    real_t a = (real_t)0.5 * rho * tworttwoGf;
    memset(H_mat, 0, sizeof(cplx) * 9);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            if (nubar == -1) H_mat[i][j] = c_rmul(-a, c_conj(mat_pot[i][j]));
            else if (nubar == 1) H_mat[i][j] = c_rmul(a, mat_pot[i][j]);
        }
}

/* numba_osc_kernels.py:571-603 get_H_decay */
void SUFFIX(oracle_get_H_decay)(const cplx mix[3][3], const cplx mix_ct[3][3],
                                const cplx mat_decay[3][3], cplx H_decay[3][3]) {
    cplx tmp[3][3];
    mat_mul(mat_decay, mix_ct, tmp);
    mat_mul(mix, tmp, H_decay);
}

/* numpy.linalg.eigvals of a general complex 3x3 (LAPACK zgeev: zgehrd + zhseqr, eigenvalues only):
 * one Givens similarity brings A to upper Hessenberg form, then shifted QR steps on the active block
 * [lo..hi] with the Wilkinson shift (the eigenvalue of the trailing 2x2 closer to its last diagonal
 * entry) until the last subdiagonal entry is negligible, deflate, repeat. */
static real_t c_abs1(cplx z) { return R_FABS(z.re) + R_FABS(z.im); }
static real_t c_abs(cplx z) { return (real_t)hypot((double)z.re, (double)z.im); }
static cplx c_sqrt(cplx z) {
    real_t m = c_abs(z);
    if (m == 0) return c_make(0, 0);
    real_t a = R_SQRT((real_t)0.5 * (m + R_FABS(z.re)));
    real_t b = (real_t)0.5 * z.im / a;
    if (z.re >= 0) return c_make(a, b);
    return c_make(R_FABS(b), z.im < 0 ? -a : a);
}
/* Givens rotation G = [[c, s], [-conj(s), c]] (c real) with G (f, g)^T = (r, 0)^T */
static void givens(cplx f, cplx g, real_t *c, cplx *s) {
    real_t ng = c_abs(g), nf = c_abs(f);
    if (ng == 0) { *c = 1; *s = c_make(0, 0); return; }
    if (nf == 0) { *c = 0; *s = c_make(1, 0); return; } /* any unit s; r = g up to phase is not needed */
    real_t nrm = (real_t)hypot((double)nf, (double)ng);
    *c = nf / nrm;
    /* s = (f / |f|) conj(g) / nrm */
    cplx fu = c_make(f.re / nf, f.im / nf);
    cplx t = c_mul(fu, c_conj(g));
    *s = c_make(t.re / nrm, t.im / nrm);
}
/* rows p, q of H <- G rows ; columns p, q of H <- columns G^H   (similarity) */
static void rot_rows(cplx H[3][3], int p, int q, real_t c, cplx s, int c0, int c1) {
    for (int j = c0; j <= c1; ++j) {
        cplx a = H[p][j], b = H[q][j];
        H[p][j] = c_add(c_rmul(c, a), c_mul(s, b));
        H[q][j] = c_sub(c_rmul(c, b), c_mul(c_conj(s), a));
    }
}
static void rot_cols(cplx H[3][3], int p, int q, real_t c, cplx s, int r0, int r1) {
    for (int i = r0; i <= r1; ++i) {
        cplx a = H[i][p], b = H[i][q];
        H[i][p] = c_add(c_rmul(c, a), c_mul(c_conj(s), b));
        H[i][q] = c_sub(c_rmul(c, b), c_mul(s, a));
    }
}
int SUFFIX(oracle_eigvals3)(const cplx A[3][3], cplx w[3]) {
    cplx H[3][3];
    mat_copy(A, H);
    const real_t eps = sizeof(real_t) == 8 ? (real_t)2.220446049250313e-16 : (real_t)1.1920929e-07;
    /* Hessenberg: annihilate H[2][0] against H[1][0] */
    {
        real_t c; cplx s;
        givens(H[1][0], H[2][0], &c, &s);
        rot_rows(H, 1, 2, c, s, 0, 2);
        rot_cols(H, 1, 2, c, s, 0, 2);
        H[2][0] = c_make(0, 0);
    }
    int hi = 2, iter = 0;
    while (hi >= 0) {
        int lo = hi;
        while (lo > 0) {
            real_t sd = c_abs1(H[lo][lo - 1]);
            real_t dd = c_abs1(H[lo][lo]) + c_abs1(H[lo - 1][lo - 1]);
            if (sd <= eps * dd || sd == 0) { H[lo][lo - 1] = c_make(0, 0); break; }
            --lo;
        }
        if (lo == hi) { w[hi] = H[hi][hi]; --hi; iter = 0; continue; }
        if (++iter > 300) return -1;
        /* Wilkinson shift from [[a, b], [c, d]] = H[hi-1..hi][hi-1..hi] */
        cplx a = H[hi - 1][hi - 1], b = H[hi - 1][hi], c_ = H[hi][hi - 1], d = H[hi][hi];
        cplx half_diff = c_rmul((real_t)0.5, c_sub(a, d));
        cplx disc = c_sqrt(c_add(c_mul(half_diff, half_diff), c_mul(b, c_)));
        cplx mean = c_rmul((real_t)0.5, c_add(a, d));
        cplx e1 = c_add(mean, disc), e2 = c_sub(mean, disc);
        cplx mu = c_abs(c_sub(e1, d)) <= c_abs(c_sub(e2, d)) ? e1 : e2;
        if (iter % 10 == 0) mu = c_add(mu, c_make(c_abs1(c_), 0)); /* exceptional shift */
        /* QR step on the active block: H - mu = QR, H <- RQ + mu */
        for (int i = lo; i <= hi; ++i) H[i][i] = c_sub(H[i][i], mu);
        real_t cs[2]; cplx ss[2];
        for (int k = lo; k < hi; ++k) {
            givens(H[k][k], H[k + 1][k], &cs[k - lo], &ss[k - lo]);
            rot_rows(H, k, k + 1, cs[k - lo], ss[k - lo], k, hi);
            H[k + 1][k] = c_make(0, 0);
        }
        for (int k = lo; k < hi; ++k) rot_cols(H, k, k + 1, cs[k - lo], ss[k - lo], lo, k + 1 < hi ? k + 2 : hi);
        for (int i = lo; i <= hi; ++i) H[i][i] = c_add(H[i][i], mu);
    }
    return 0;
}

/* numba_osc_kernels.py:656-685 get_dms_numerical */
int SUFFIX(oracle_get_dms_numerical)(real_t energy, const cplx H[3][3], cplx dm_mat_mat[3][3],
                                     cplx dm_mat[3][3]) {
    cplx w[3], m[3];
    if (SUFFIX(oracle_eigvals3)(H, w)) return -1;
    for (int i = 0; i < 3; ++i) m[i] = c_rmul((real_t)2.0 * energy, w[i]);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            dm_mat_mat[i][j] = c_sub(m[i], m[j]);
            dm_mat[i][j] = m[i];
        }
    return 0;
}

/* numba_osc_kernels.py:687-831 get_dms */
void SUFFIX(oracle_get_dms)(real_t energy, const cplx H[3][3], const real_t dmv[3][3],
                            cplx dm_mat_mat[3][3], cplx dm_mat[3][3]) {
    real_t real_product_a = c_mul(c_mul(H[0][1], H[1][2]), H[2][0]).re;
    real_t real_product_b = c_mul(c_mul(H[0][0], H[1][1]), H[2][2]).re;

    real_t n_emu = H[0][1].re * H[0][1].re + H[0][1].im * H[0][1].im;
    real_t n_etau = H[0][2].re * H[0][2].re + H[0][2].im * H[0][2].im;
    real_t n_mutau = H[1][2].re * H[1][2].re + H[1][2].im * H[1][2].im;

    cplx s12 = c_add(H[1][1], H[2][2]);
    real_t c1 = c_rmul(H[0][0].re, s12).re - c_rmul(H[0][0].im, s12).im +
                c_rmul(H[1][1].re, H[2][2]).re - c_rmul(H[1][1].im, H[2][2]).im -
                n_emu - n_mutau - n_etau;

    real_t c0 = H[0][0].re * n_mutau + H[1][1].re * n_etau + H[2][2].re * n_emu -
                (real_t)2.0 * real_product_a - real_product_b;

    real_t c2 = -H[0][0].re - H[1][1].re - H[2][2].re;

    real_t one_over_two_e = (real_t)0.5 / energy;
    real_t one_third = (real_t)(1.0 / 3.0);
    real_t two_third = (real_t)(2.0 / 3.0);

    real_t x = dmv[1][0];
    real_t y = dmv[2][0];

    real_t c2_v = -one_over_two_e * (x + y);

    real_t p = c2 * c2 - (real_t)3.0 * c1;
    real_t p_v = one_over_two_e * one_over_two_e * (x * x + y * y - x * y);
    p = p > 0 ? p : (real_t)0; /* max(0.0, p); nan -> 0.0 like Python's max(0.0, nan) */

    real_t q = (real_t)-13.5 * c0 - c2 * c2 * c2 + (real_t)4.5 * c1 * c2;
    real_t q_v = one_over_two_e * one_over_two_e * one_over_two_e * (x + y) *
                 ((x + y) * (x + y) - (real_t)4.5 * x * y);

    real_t tmp = (real_t)27 * ((real_t)0.25 * (c1 * c1) * (p - c1) + c0 * (q + (real_t)6.75 * c0));
    real_t tmp_v = p_v * p_v * p_v - q_v * q_v;
    tmp = tmp > 0 ? tmp : (real_t)0;

    real_t theta[3], theta_v[3], m_mat[3], m_mat_u[3], m_mat_v[3];
    real_t a = two_third * (real_t)M_PI;
    real_t res = R_ATAN2(R_SQRT(tmp), q) * one_third;
    theta[0] = res + a; theta[1] = res - a; theta[2] = res;
    real_t res_v = R_ATAN2(R_SQRT(tmp_v), q_v) * one_third;
    theta_v[0] = res_v + a; theta_v[1] = res_v - a; theta_v[2] = res_v;

    real_t b = two_third * R_SQRT(p);
    real_t b_v = two_third * R_SQRT(p_v);

    for (int i = 0; i < 3; ++i) {
        m_mat_u[i] = (real_t)2.0 * energy * (b * R_COS(theta[i]) - c2 * one_third + dmv[0][0]);
        m_mat_v[i] = (real_t)2.0 * energy * (b_v * R_COS(theta_v[i]) - c2_v * one_third + dmv[0][0]);
    }

    /* sort according to which reproduce the vacuum eigenstates (:816-825) */
    for (int i = 0; i < 3; ++i) {
        real_t best = R_FABS(dmv[i][0] - m_mat_v[0]);
        int k = 0;
        for (int j = 0; j < 3; ++j) {
            real_t t = R_FABS(dmv[i][0] - m_mat_v[j]);
            if (t < best) { k = j; best = t; }
        }
        m_mat[i] = m_mat_u[k];
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            dm_mat_mat[i][j] = c_make(m_mat[i] - m_mat[j], 0);
            dm_mat[i][j] = c_make(m_mat[i], 0);
        }
}

/* numba_osc_kernels.py:834-872 get_product; product is [3][3][3] */
void SUFFIX(oracle_get_product)(real_t energy, const cplx dm_mat[3][3], const cplx dm_mat_mat[3][3],
                                const cplx Hm[3][3], cplx product[3][3][3]) {
    cplx HmM[3][3][3];
    real_t two_e = (real_t)2.0 * energy;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            for (int k = 0; k < 3; ++k) {
                HmM[i][j][k] = c_rmul(two_e, Hm[i][j]);
                if (i == j) HmM[i][j][k] = c_sub(HmM[i][j][k], dm_mat[k][j]);
                product[i][j][k] = c_make(0, 0);
            }
    /* denominators: complex products (real-valued without decay, where c_div is the per-component division) */
    cplx d0 = c_mul(dm_mat_mat[0][1], dm_mat_mat[0][2]);
    cplx d1 = c_mul(dm_mat_mat[1][2], dm_mat_mat[1][0]);
    cplx d2 = c_mul(dm_mat_mat[2][0], dm_mat_mat[2][1]);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            for (int k = 0; k < 3; ++k) {
                product[i][j][0] = c_add(product[i][j][0], c_mul(HmM[i][k][1], HmM[k][j][2]));
                product[i][j][1] = c_add(product[i][j][1], c_mul(HmM[i][k][2], HmM[k][j][0]));
                product[i][j][2] = c_add(product[i][j][2], c_mul(HmM[i][k][0], HmM[k][j][1]));
            }
            product[i][j][0] = c_div(product[i][j][0], d0);
            product[i][j][1] = c_div(product[i][j][1], d1);
            product[i][j][2] = c_div(product[i][j][2], d2);
        }
}

/* numba_osc_kernels.py:481-531 get_transition_matrix_massbasis */
void SUFFIX(oracle_get_transition_matrix_massbasis)(real_t baseline, real_t energy,
                                                    const cplx dm_mat[3][3],
                                                    const cplx dm_mat_mat[3][3],
                                                    const cplx Hm[3][3], cplx T[3][3]) {
    cplx product[3][3][3];
    memset(T, 0, sizeof(cplx) * 9);
    SUFFIX(oracle_get_product)(energy, dm_mat, dm_mat_mat, Hm, product);
    const real_t hbar_c_factor = (real_t)2.534;
    for (int k = 0; k < 3; ++k) {
        /* arg = -dm_mat[k,0] * (baseline / energy) * hbar_c_factor is complex (imaginary part 0 without decay);
         * c = cmath.exp(arg * 1j) = exp(-Im arg) (cos(Re arg) + i sin(Re arg)) */
        real_t boe = baseline / energy;
        real_t arg = -dm_mat[k][0].re * boe * hbar_c_factor;
        real_t arg_im = -dm_mat[k][0].im * boe * hbar_c_factor;
        cplx c = c_make(R_COS(arg), R_SIN(arg));
        if (arg_im != 0) { real_t damp = (real_t)exp((double)-arg_im); c = c_make(damp * c.re, damp * c.im); }
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) T[i][j] = c_add(T[i][j], c_mul(c, product[i][j][k]));
    }
}

/* numba_osc_kernels.py:348-478 get_transition_matrix */
int SUFFIX(oracle_get_transition_matrix)(int64_t nubar, real_t energy, real_t rho, real_t baseline,
                                         const cplx mix_nubar[3][3], const cplx mix_nubar_ct[3][3],
                                         const cplx mat_pot[3][3], const cplx H_vac[3][3],
                                         int64_t decay_flag, const cplx H_decay[3][3],
                                         const real_t lri_pot[3][3], const real_t dm[3][3],
                                         cplx T[3][3]) {
    cplx H_mat[3][3], dm_mat[3][3], dm_mat_mat[3][3], H_full[3][3], tmp[3][3], Hm[3][3];
    SUFFIX(oracle_get_H_mat)(rho, mat_pot, nubar, H_mat);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            if (nubar > 0) H_mat[i][j].re = H_mat[i][j].re + lri_pot[i][j] * (real_t)1e9;
            else if (nubar < 0) H_mat[i][j].re = H_mat[i][j].re - lri_pot[i][j] * (real_t)1e9;
        }
    real_t one_over_two_e = (real_t)0.5 / energy;
    if (decay_flag == 1) { /* :445-451 */
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                H_full[i][j] = c_add(c_mul(c_add(H_vac[i][j], H_decay[i][j]), c_make(one_over_two_e, 0)), H_mat[i][j]);
        if (SUFFIX(oracle_get_dms_numerical)(energy, H_full, dm_mat_mat, dm_mat)) return -4;
    } else {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                H_full[i][j] = c_add(c_mul(H_vac[i][j], c_make(one_over_two_e, 0)), H_mat[i][j]);
        SUFFIX(oracle_get_dms)(energy, H_full, dm, dm_mat_mat, dm_mat);
    }
    mat_mul(H_full, mix_nubar, tmp);
    mat_mul(mix_nubar_ct, tmp, Hm);
    SUFFIX(oracle_get_transition_matrix_massbasis)(baseline, energy, dm_mat, dm_mat_mat, Hm, T);
    return 0;
}

/* numba_osc_kernels.py:121-345 osc_probs_layers_kernel (cache = True branch) */
int SUFFIX(oracle_osc_probs_layers)(const real_t dm[3][3], const cplx mix[3][3],
                                    const cplx mat_pot[3][3], int64_t decay_flag,
                                    const cplx mat_decay[3][3], const real_t lri_pot[3][3],
                                    int64_t nubar, real_t energy, const real_t *density,
                                    const real_t *distance, int n_layers, real_t osc_probs[3][3]) {
    cplx H_vac[3][3], H_decay[3][3], mixn[3][3], mixn_ct[3][3], prod[3][3], T[3][3], tmp[3][3];
    /* transition_matrices has 120 slots (:227) but the loops run over the array's width (:230,282): PREM_59layer
     * arrives 122 wide with at most 118 active slots and works; an ACTIVE slot >= 120 would be written out of bounds */
    for (int i = MAX_LAYERS; i < n_layers; ++i)
        if (distance[i] > 0) return -3;
    if (n_layers > MAX_LAYERS) n_layers = MAX_LAYERS;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            mixn[i][j] = nubar > 0 ? mix[i][j] : c_conj(mix[i][j]);
            osc_probs[i][j] = 0;
        }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) mixn_ct[j][i] = c_conj(mixn[i][j]);
    SUFFIX(oracle_get_H_vac)(mixn, mixn_ct, dm, H_vac);
    SUFFIX(oracle_get_H_decay)(mixn, mixn_ct, mat_decay, H_decay); /* :220 (only read when decay_flag == 1) */

    cplx Ts[MAX_LAYERS][3][3];
    for (int i = 0; i < n_layers; ++i) {
        real_t rho = density[i], d = distance[i];
        if (d > 0) {
            int hit = -1;
            for (int j = 0; j < i; ++j)
                if (R_FABS(density[j] - rho) < (real_t)1e-5 && R_FABS(distance[j] - d) < (real_t)1e-5)
                    hit = j; /* last match wins (:236-241) */
            if (hit >= 0) mat_copy(Ts[hit], Ts[i]);
            else {
                int rc = SUFFIX(oracle_get_transition_matrix)(nubar, energy, rho, d, mixn, mixn_ct,
                                                              mat_pot, H_vac, decay_flag, H_decay,
                                                              lri_pot, dm, T);
                if (rc) return rc;
                mat_copy(T, Ts[i]);
            }
        }
    }
    int first = 1;
    memset(prod, 0, sizeof prod);
    for (int i = 0; i < n_layers; ++i) {
        if (distance[i] > 0) {
            if (first) { mat_copy(Ts[i], prod); first = 0; }
            else { mat_mul(Ts[i], prod, tmp); mat_copy(tmp, prod); }
        }
    }
    mat_mul(prod, mixn_ct, tmp);
    mat_mul(mixn, tmp, prod);
    for (int i = 0; i < 3; ++i) {
        /* output_psi = prod . e_i  (matrix_dot_vector with explicit zero terms) */
        for (int r = 0; r < 3; ++r) {
            cplx acc = c_make(0, 0);
            for (int n = 0; n < 3; ++n)
                acc = c_add(acc, c_mul(prod[r][n], c_make(n == i ? (real_t)1 : (real_t)0, 0)));
            osc_probs[i][r] += acc.re * acc.re + acc.im * acc.im;
        }
    }
    return 0;
}

/* numba_osc_hostfuncs.py:60-70 propagate_array: one kernel call per event.
 * nubar is either a scalar broadcast (nubar_stride = 0) or per event.
 * OpenMP over events stands in for numba's target='parallel'. */
int SUFFIX(oracle_propagate_array)(const real_t *dm, const real_t *mix_ri, const real_t *mat_pot_ri,
                                   int64_t decay_flag, const real_t *mat_decay_ri,
                                   const real_t *lri_pot, const int64_t *nubar, int nubar_stride,
                                   const real_t *energy, const real_t *densities,
                                   const real_t *distances, int64_t n_events, int n_layers,
                                   real_t *probability, int n_threads) {
    int rc_all = 0;
    (void)n_threads;
#pragma omp parallel for schedule(dynamic, 256) num_threads(n_threads) reduction(|:rc_all)
    for (int64_t e = 0; e < n_events; ++e) {
        int rc = SUFFIX(oracle_osc_probs_layers)(
            (const real_t(*)[3])dm, (const cplx(*)[3])mix_ri, (const cplx(*)[3])mat_pot_ri,
            decay_flag, (const cplx(*)[3])mat_decay_ri, (const real_t(*)[3])lri_pot,
            nubar[(size_t)e * nubar_stride], energy[e], densities + (size_t)e * n_layers,
            distances + (size_t)e * n_layers, n_layers, (real_t(*)[3])(probability + (size_t)e * 9));
        rc_all |= (rc != 0);
    }
    return rc_all ? -1 : 0;
}

/* numba_osc_hostfuncs.py:206-221 fill_probs */
void SUFFIX(oracle_fill_probs)(const real_t *probability, int initial_flav, int flav,
                               int64_t n_events, real_t *out) {
    for (int64_t e = 0; e < n_events; ++e) out[e] = probability[(size_t)e * 9 + initial_flav * 3 + flav];
}

/* prob3.py:621-622 apply_function: weights *= flux_e*prob_e + flux_mu*prob_mu */
void SUFFIX(oracle_apply_osc_weights)(const real_t *nu_flux /* [N,2] */, const real_t *prob_e,
                                      const real_t *prob_mu, int64_t n_events, real_t *weights) {
    for (int64_t e = 0; e < n_events; ++e)
        weights[e] *= (nu_flux[2 * e] * prob_e[e]) + (nu_flux[2 * e + 1] * prob_mu[e]);
}
