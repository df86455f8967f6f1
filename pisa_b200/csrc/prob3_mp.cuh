// prob3_mp.cuh -- the FP32 mode of the propagation: mixed-precision arithmetic.
//
// The reference's FP32 mode (PISA_FTYPE=fp32, pisa/__init__.py:152-179) runs the whole kernel in float32 /
// complex64.  A plain float port cannot keep the BASELINE tolerance (1e-5 absolute on probabilities up to 1 TeV):
//   * oscillation phases (lambda_i - lambda_j) * t reach ~1e2 rad, so eigenvalue DIFFERENCES need ~1e-9 relative;
//   * the Lagrange / Cayley-Hamilton coefficients cancel by 1 / (gap * t) when two eigenvalues are close (every
//     event above ~50 GeV in matter), which amplifies any float error in sin / cos by 1e3 .. 1e5.
// What is measured on B200 (scratch/micro/pipe_mix.cu, profiles/r02_pipe_mix.txt): DFMA issues once per 2 cycles per
// sub-partition, scalar FFMA once per cycle, and the two pipes run CONCURRENTLY (DFMA + FFMA interleaved: 0.92
// instructions / cycle).  The FP64 kernel is bound by the FP64 pipe (1319 of 2316 instructions per event).  So the
// FP32 mode keeps in FP64 only what needs it and moves everything else to the FP32 pipe:
//   FP64  characteristic cubic (3 FMAs per layer from per-event invariants), its roots (eigen_roots_centered), the two
//         phase arguments gap * t and their reduction modulo pi/2, the shell geometry (segment lengths);
//   FP32  sin / cos polynomials on the reduced arguments, returned as E = exp(-i delta) - 1 with RELATIVE accuracy
//         for small delta; Newton (divided-difference) coefficients
//             n1 = E_ab / g_ba ,  n2 = (E_ac / g_ca - E_ab / g_ba) / g_cb ,
//         which need no cancellation-prone Lagrange weights; the transition matrix
//             T = 1 + n1 (M - mu_a) + n2 (M - mu_a)(M - mu_b)       (global phase exp(-i mu_a t) dropped)
//         on the trace-free M = H - tr(H)/3, whose off-diagonal products are per-event invariants when only the
//         matter term moves (standard matter potential); all matrix-vector products; the propagation state, which
//         now fits in registers (18 floats) instead of shared memory.
// Error budget: every float quantity carries 6e-8 relative, the matrices multiply phases up to ~50 rad, so a layer
// contributes <~ 3e-6 in the worst case; measured against the FP64 oracle on float32-rounded inputs:
// tests/test_device_math_emulation.py (host emulation) and tests/test_gpu_prob3.py (GPU).
#pragma once
#include "prob3_device.cuh"

namespace pisab {

// Two arithmetic "real" types share the float part of this file:
//   float : one event per thread;
//   f2    : TWO events per thread, one in each lane of the packed FP32 instructions of sm_100 (fma.rn.f32x2 ...).
//           A packed instruction occupies the FP32 pipe for two cycles -- the same pipe time per event as two
//           scalar ones -- but only ONE issue slot, and the FP32 mode is issue-bound; ptxas folds negations into the
//           operand modifiers of FFMA2 / FADD2 / FMUL2, so every scalar float operation below has a one-instruction
//           packed twin and both forms round identically (the pair path is bit-identical to the scalar one).
struct alignas(8) f2 {
    float x, y; // lane x: first event of the pair, lane y: second
};
#ifdef PISAB_HOST_EMU
__device__ __forceinline__ f2 f2_fma(f2 a, f2 b, f2 c) { return f2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
__device__ __forceinline__ f2 f2_mul(f2 a, f2 b) { return f2{a.x * b.x, a.y * b.y}; }
__device__ __forceinline__ f2 f2_add(f2 a, f2 b) { return f2{a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ f2 f2_sub(f2 a, f2 b) { return f2{a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ f2 f2_neg(f2 a) { return f2{-a.x, -a.y}; }
#else
__device__ __forceinline__ unsigned long long f2_bits(f2 a) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
    return r;
}
__device__ __forceinline__ f2 f2_from(unsigned long long v) {
    f2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ f2 f2_fma(f2 a, f2 b, f2 c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)), "l"(f2_bits(c)));
    return f2_from(r);
}
__device__ __forceinline__ f2 f2_mul(f2 a, f2 b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return f2_from(r);
}
__device__ __forceinline__ f2 f2_add(f2 a, f2 b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return f2_from(r);
}
__device__ __forceinline__ f2 f2_sub(f2 a, f2 b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return f2_from(r);
}
__device__ __forceinline__ f2 f2_neg(f2 a) { return f2{-a.x, -a.y}; } // folded into the consumer's operand modifier
#endif
// the common vocabulary of the templated code: t_fma(a, b, c) = a b + c, t_mul, t_add, t_sub, t_neg, t_splat<R>(float)
__device__ __forceinline__ float t_fma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ float t_mul(float a, float b) { return a * b; }
__device__ __forceinline__ float t_add(float a, float b) { return a + b; }
__device__ __forceinline__ float t_sub(float a, float b) { return a - b; }
__device__ __forceinline__ float t_neg(float a) { return -a; }
__device__ __forceinline__ f2 t_fma(f2 a, f2 b, f2 c) { return f2_fma(a, b, c); }
__device__ __forceinline__ f2 t_mul(f2 a, f2 b) { return f2_mul(a, b); }
__device__ __forceinline__ f2 t_add(f2 a, f2 b) { return f2_add(a, b); }
__device__ __forceinline__ f2 t_sub(f2 a, f2 b) { return f2_sub(a, b); }
__device__ __forceinline__ f2 t_neg(f2 a) { return f2_neg(a); }
template <typename R> __device__ __forceinline__ R t_splat(float v);
template <> __device__ __forceinline__ float t_splat<float>(float v) { return v; }
template <> __device__ __forceinline__ f2 t_splat<f2>(float v) { return f2{v, v}; }

template <typename R>
struct CplxT {
    R re, im;
};
typedef CplxT<float> CplxF;
typedef CplxT<f2> Cplx2;
typedef CplxF Mat3F[3][3];

template <typename R>
__device__ __forceinline__ CplxT<R> t_cmul(CplxT<R> a, CplxT<R> b) {
    return CplxT<R>{t_fma(a.re, b.re, t_neg(t_mul(a.im, b.im))), t_fma(a.re, b.im, t_mul(a.im, b.re))};
}
template <typename R>
__device__ __forceinline__ CplxT<R> t_cfma(CplxT<R> a, CplxT<R> b, CplxT<R> c) { // a*b + c
    CplxT<R> r;
    r.re = t_fma(a.re, b.re, t_fma(t_neg(a.im), b.im, c.re));
    r.im = t_fma(a.re, b.im, t_fma(a.im, b.re, c.im));
    return r;
}

__device__ __forceinline__ CplxF cmulf(CplxF a, CplxF b) {
    return CplxF{fmaf(a.re, b.re, -a.im * b.im), fmaf(a.re, b.im, a.im * b.re)};
}
__device__ __forceinline__ CplxF cfmaf(CplxF a, CplxF b, CplxF c) { // a*b + c
    CplxF r;
    r.re = fmaf(a.re, b.re, fmaf(-a.im, b.im, c.re));
    r.im = fmaf(a.re, b.im, fmaf(a.im, b.re, c.im));
    return r;
}

__device__ __forceinline__ float rcp_f32(float x) {
#ifdef PISAB_HOST_EMU
    return 1.0f / x;
#else
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}

// E = exp(-i delta) - 1 = (cos(delta) - 1, -sin(delta)).  The argument is reduced in FP64 (delta reaches ~1e2 rad and
// must keep ~1e-8 absolute), the polynomials run in float on |r| <= pi/4 (Cephes sinf / cosf kernels, ~1 ulp).  For
// |delta| < pi/4 (quadrant 0) both components keep RELATIVE accuracy -- cos - 1 is evaluated without the leading 1 --
// which is what makes the divided differences below safe when two eigenvalues are close.
__device__ __forceinline__ CplxF expm1i_neg(double delta) {
    const double shifted = fma(delta, kTab[0], kTab[21]);
    const int k = __double2loint(shifted);
    const double kd = shifted - kTab[21];
    const float r = (float)fma(-kd, kTab[1], delta); // (the pi/2 low word is < 4e-15 here: irrelevant for float)
    const float z = r * r;
    float ps = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = fmaf(z, ps, -1.6666654611e-1f);
    const float sn = fmaf(r * z, ps, r);
    float pc = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = fmaf(z, pc, 4.166664568298827e-2f);
    pc = fmaf(z, pc, -0.5f);
    const float cm = z * pc; // cos r - 1
    //  q = k mod 4:   cos(delta) - 1 =  cm | -sn - 1 | -cm - 2 |  sn - 1      sin(delta) = sn | 1 + cm | -sn | -(1 + cm)
    const bool odd = k & 1;
    float c1 = odd ? sn : cm;
    c1 = ((k + 1) & 2) ? -c1 : c1;
    const float off = (k & 3) == 0 ? 0.0f : ((k & 3) == 2 ? -2.0f : -1.0f);
    float sd = odd ? 1.0f + cm : sn;
    sd = (k & 2) ? -sd : sd;
    return CplxF{c1 + off, -sd};
}

// Centered roots (ascending, sum 0) of the characteristic cubic x^3 + c1 x + c0 of a TRACE-FREE Hermitian 3x3:
// the trigonometric solution of numba_osc_kernels.py:766-814 with c2 = 0, in FP64 (see eigen_solve).
struct RootsC {
    double m0, m1, m2;
};
__device__ __forceinline__ RootsC eigen_roots_centered(double c1, double c0) {
    double p = -3.0 * c1;
    p = p > kTab[20] ? p : kTab[20];
    const double q = -13.5 * c0;
    const double disc = 27.0 * fma(0.25 * c1 * c1, p - c1, c0 * fma(6.75, c0, q));
    const double rs = rsqrt_fast(p);
    const double b = kTab[17] * (p * rs);
    const double inv = rs * rs * rs;
    double st, ct;
    unit_cube_root_mp(q * inv, sqrt_pos(disc) * inv, &ct, &st);
    const double kh = 0.5, ks = kTab[15];
    RootsC r;
    r.m0 = b * (-kh * ct - ks * st);
    r.m1 = b * (-kh * ct + ks * st);
    r.m2 = b * ct;
    return r;
}

// The float image of one layer's trace-free Hamiltonian M (off-diagonal part) and the invariants its square needs:
//   s_i  = sum_{j != i} |h_ij|^2                      (diagonal of M^2 minus m_i^2)
//   c_ij = h_ik h_kj, k the third index               (off-diagonal of M^2 minus (m_i + m_j) h_ij)
// R = float: one event; R = f2: the two events of a pair, lane-wise.
template <typename R>
struct LayerMatT {
    R r01, i01, r02, i02, r12, i12;
    R s0, s1, s2;
    CplxT<R> c01, c02, c12;
};
typedef LayerMatT<float> LayerMatF;
template <typename R>
__device__ __forceinline__ void layer_products(LayerMatT<R> &L) {
    const R n01 = t_fma(L.r01, L.r01, t_mul(L.i01, L.i01));
    const R n02 = t_fma(L.r02, L.r02, t_mul(L.i02, L.i02));
    const R n12 = t_fma(L.r12, L.r12, t_mul(L.i12, L.i12));
    L.s0 = t_add(n01, n02);
    L.s1 = t_add(n01, n12);
    L.s2 = t_add(n02, n12);
    L.c01 = CplxT<R>{t_fma(L.r02, L.r12, t_mul(L.i02, L.i12)), t_fma(L.i02, L.r12, t_neg(t_mul(L.r02, L.i12)))};  // h02 conj(h12)
    L.c02 = CplxT<R>{t_fma(L.r01, L.r12, t_neg(t_mul(L.i01, L.i12))), t_fma(L.r01, L.i12, t_mul(L.i01, L.r12))};  // h01 h12
    L.c12 = CplxT<R>{t_fma(L.r01, L.r02, t_mul(L.i01, L.i02)), t_fma(L.r01, L.i02, t_neg(t_mul(L.i01, L.r02)))};  // conj(h01) h02
}

// Scalar coefficients of one layer of one event:  T = 1 + n1 (M - mu0) + n2 (M - mu0)(M - mu1)  (global phase
// exp(+i mu0 t) dropped), with the diagonal of M given through ea_i = m_i - mu0 (formed in FP64 by the caller and
// rounded once); m_i - mu1 = ea_i - g10 and mu2 - m_k = g20 - ea_k.
template <typename R>
struct LayerCoefT {
    CplxT<R> n1, n2;
    R g10, g20, ea0, ea1, ea2;
};
typedef LayerCoefT<float> LayerCoefF;
// FP64 roots -> phases (FP64 reduction) -> float divided differences; per event, always scalar
__device__ __forceinline__ LayerCoefF layer_coefficients(float ea0, float ea1, float ea2, const RootsC &R, double t) {
    const double g10d = R.m1 - R.m0, g20d = R.m2 - R.m0, g21d = R.m2 - R.m1;
    const CplxF e01 = expm1i_neg(g10d * t), e02 = expm1i_neg(g20d * t);
    const float g10 = (float)g10d, g20 = (float)g20d, g21 = (float)g21d;
    // (a gap that underflows in float would give 0 * inf: clamp; three equal roots cannot occur, see eigen_solve)
    const float r10 = rcp_f32(fmaxf(g10, 1e-30f)), r20 = rcp_f32(g20), r21 = rcp_f32(fmaxf(g21, 1e-30f));
    LayerCoefF k;
    k.n1 = CplxF{e01.re * r10, e01.im * r10};
    const CplxF f02{e02.re * r20, e02.im * r20};
    k.n2 = CplxF{(f02.re - k.n1.re) * r21, (f02.im - k.n1.im) * r21};
    k.g10 = g10; k.g20 = g20; k.ea0 = ea0; k.ea1 = ea1; k.ea2 = ea2;
    return k;
}
__device__ __forceinline__ LayerCoefT<f2> pack_coef(const LayerCoefF &a, const LayerCoefF &b) {
    LayerCoefT<f2> k;
    k.n1 = Cplx2{f2{a.n1.re, b.n1.re}, f2{a.n1.im, b.n1.im}};
    k.n2 = Cplx2{f2{a.n2.re, b.n2.re}, f2{a.n2.im, b.n2.im}};
    k.g10 = f2{a.g10, b.g10}; k.g20 = f2{a.g20, b.g20};
    k.ea0 = f2{a.ea0, b.ea0}; k.ea1 = f2{a.ea1, b.ea1}; k.ea2 = f2{a.ea2, b.ea2};
    return k;
}

//   (M - mu0)(M - mu1)_ii = (m_i - mu0)(m_i - mu1) + s_i
//   (M - mu0)(M - mu1)_ij = h_ij (mu2 - m_k) + c_ij          (trace-free: m_i + m_j = -m_k, mu0 + mu1 = -mu2)
template <typename R>
__device__ __forceinline__ void assemble_matrix(const LayerMatT<R> &L, const LayerCoefT<R> &k, CplxT<R> (*T)[3]) {
    const CplxT<R> n1 = k.n1, n2 = k.n2;
    const R one = t_splat<R>(1.0f);
#define PISAB_MP_DIAG(I, EA, S)                                                                          \
    {                                                                                                    \
        const R pp = t_fma(EA, t_sub(EA, k.g10), S);                                                     \
        T[I][I] = CplxT<R>{t_fma(n2.re, pp, t_fma(n1.re, EA, one)), t_fma(n2.im, pp, t_mul(n1.im, EA))}; \
    }
    PISAB_MP_DIAG(0, k.ea0, L.s0)
    PISAB_MP_DIAG(1, k.ea1, L.s1)
    PISAB_MP_DIAG(2, k.ea2, L.s2)
#undef PISAB_MP_DIAG
    // off-diagonal pairs: T_ij = h z + n2 c, T_ji = conj(h) z + n2 conj(c), z = n1 + n2 (mu2 - m_k)
#define PISAB_MP_OFF(I, J, EAK, HR, HI, C)                                         \
    {                                                                              \
        const R u = t_sub(k.g20, EAK);                                             \
        const R zr = t_fma(n2.re, u, n1.re), zi = t_fma(n2.im, u, n1.im);          \
        const R s1 = t_fma(HR, zr, t_mul(n2.re, C.re));                            \
        const R s2 = t_fma(HI, zi, t_mul(n2.im, C.im));                            \
        const R s3 = t_fma(HR, zi, t_mul(n2.im, C.re));                            \
        const R s4 = t_fma(HI, zr, t_mul(n2.re, C.im));                            \
        T[I][J] = CplxT<R>{t_sub(s1, s2), t_add(s3, s4)};                          \
        T[J][I] = CplxT<R>{t_add(s1, s2), t_sub(s3, s4)};                          \
    }
    PISAB_MP_OFF(0, 1, k.ea2, L.r01, L.i01, L.c01)
    PISAB_MP_OFF(0, 2, k.ea1, L.r02, L.i02, L.c02)
    PISAB_MP_OFF(1, 2, k.ea0, L.r12, L.i12, L.c12)
#undef PISAB_MP_OFF
}

// Propagation state of the FP32 mode in a per-thread column of shared memory, [(v*3+k)][thread]: one (re, im) float2
// per complex number (R = float, 64-bit accesses) or one (re_A, re_B, im_A, im_B) float4 for a pair (R = f2, 128-bit
// accesses).  A register-resident state costs ~190 loop-carried register moves per event (ncu, round 2), this form
// ~70 conflict-free LDS / STS.  Vectors 0..NC-1 = columns of R, NC.. = rows of L.
template <typename R> struct StateSlot;
template <> struct StateSlot<float> {
    typedef float2 type;
    static __device__ __forceinline__ CplxF get(const float2 &z) { return CplxF{z.x, z.y}; }
    static __device__ __forceinline__ float2 put(const CplxF &z) { return make_float2(z.re, z.im); }
};
template <> struct StateSlot<f2> {
    typedef float4 type;
    static __device__ __forceinline__ Cplx2 get(const float4 &z) { return Cplx2{f2{z.x, z.y}, f2{z.z, z.w}}; }
    static __device__ __forceinline__ float4 put(const Cplx2 &z) { return make_float4(z.re.x, z.re.y, z.im.x, z.im.y); }
};
template <int NR, int NC, typename R>
struct PropagatorSmemT {
    static constexpr bool kF32 = true;
    typedef CplxT<R> cplx;
    typedef typename StateSlot<R>::type slot;
    slot *col; // &state[0][threadIdx.x]
    int pitch; // block size
    static constexpr int kSlots = (NR + NC) * 3;

    __device__ __forceinline__ cplx ld(int v, int k) const { return StateSlot<R>::get(col[(v * 3 + k) * pitch]); }
    __device__ __forceinline__ void st(int v, int k, cplx z) { col[(v * 3 + k) * pitch] = StateSlot<R>::put(z); }
    __device__ __forceinline__ void set_right(int c, int k, cplx v) { st(c, k, v); }
    __device__ __forceinline__ void init_right(const cplx (*T)[3]) {
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int k = 0; k < 3; ++k) st(c, k, T[k][c]);
    }
    __device__ __forceinline__ void init_left(const cplx (*T)[3], int flav) {
        if (NR == 3) {
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) st(NC + r, c, T[r][c]);
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                cplx v = T[0][c];
                if (flav == 1) v = T[1][c];
                if (flav == 2) v = T[2][c];
                st(NC, c, v);
            }
        }
    }
    __device__ __forceinline__ void mul_right(const cplx (*T)[3]) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const cplx r0 = ld(c, 0), r1 = ld(c, 1), r2 = ld(c, 2);
#pragma unroll
            for (int k = 0; k < 3; ++k) st(c, k, t_cfma(T[k][2], r2, t_cfma(T[k][1], r1, t_cmul(T[k][0], r0))));
        }
    }
    __device__ __forceinline__ void mul_left(const cplx (*T)[3]) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const cplx l0 = ld(NC + r, 0), l1 = ld(NC + r, 1), l2 = ld(NC + r, 2);
#pragma unroll
            for (int c = 0; c < 3; ++c) st(NC + r, c, t_cfma(l2, T[2][c], t_cfma(l1, T[1][c], t_cmul(l0, T[0][c]))));
        }
    }
    // |amplitude|^2 of (row r, column c); one value per lane
    __device__ __forceinline__ R prob_r(int r, int c) const {
        const cplx acc = t_cfma(ld(NC + r, 2), ld(c, 2), t_cfma(ld(NC + r, 1), ld(c, 1), t_cmul(ld(NC + r, 0), ld(c, 0))));
        return t_fma(acc.re, acc.re, t_mul(acc.im, acc.im));
    }
    __device__ __forceinline__ double prob(int r, int c) const; // scalar form only
};
template <int NR, int NC>
using PropagatorSmemF = PropagatorSmemT<NR, NC, float>;
template <int NR, int NC>
using PropagatorSmemP = PropagatorSmemT<NR, NC, f2>;
template <int NR, int NC, typename R>
__device__ __forceinline__ double PropagatorSmemT<NR, NC, R>::prob(int r, int c) const {
    static_assert(sizeof(R) == sizeof(float), "prob() is the one-event form; use prob_r() for a pair");
    return (double)prob_r(r, c);
}

__device__ __forceinline__ void herm_offdiag_to_float(const Herm3 &c, LayerMatF &L) {
    L.r01 = (float)c.r01; L.i01 = (float)c.i01; L.r02 = (float)c.r02;
    L.i02 = (float)c.i02; L.r12 = (float)c.r12; L.i12 = (float)c.i12;
}

// Per-event FP64 side of the FP32 mode (everything that needs double precision lives here):
//   STD (standard matter potential, vm = diag(a, 0, 0)): per event the trace-free H0c = hv/E + lr - tr/3 and the
//   invariants that make the cubic of H0c + x e00 - x/3 a polynomial in x = rho a;
//   general (NSI): the layer Hamiltonian is formed and centred in FP64 per layer.
template <bool STD>
struct H0Core {
    Herm3 h;                               // general: H0 = hv/E + lr (STD: only used by init)
    double c1_0, c0_0, d0c, d1c, d2c, m00; // STD

    // returns the trace-free H0c (STD) whose off-diagonal part the caller rounds to float
    __device__ __forceinline__ Herm3 init(const Herm3 &h0) {
        Herm3 c = h0;
        if (STD) {
            const double tr3 = (h0.d0 + h0.d1 + h0.d2) * kTab[16];
            c.d0 -= tr3; c.d1 -= tr3; c.d2 -= tr3;
            double c2;
            char_poly(c, c2, c1_0, c0_0); // c2 == 0 up to rounding
            d0c = c.d0; d1c = c.d1; d2c = c.d2;
            m00 = fma(c.d1, c.d2, -fma(c.r12, c.r12, c.i12 * c.i12));
        } else {
            h = h0;
        }
        return c;
    }
    // coefficients of a layer of density rho and t = 2 * 2.534 * length; general path: also the layer's own float
    // off-diagonal part in `Lgen`
    __device__ __forceinline__ LayerCoefF layer(double rho, const Herm3 &vm, double t, LayerMatF *Lgen) const {
        if (STD) {
            // H = H0c + x e00 has trace x; M = H - x/3 is trace-free with diagonal shifts (2x/3, -x/3, -x/3) and
            // det(M - mu) from the cubic of H0c + x e00 (c2 = -x, c1 = c1_0 - x d0c, c0 = c0_0 - x m00) shifted
            // by x/3:  with y = x/3:  c1' = c1 - 3 y^2 ,  c0' = c0 + y c1 - 2 y^3   (c2 = -3y)
            const double x = rho * vm.d0, y = x * kTab[16];
            const double c1 = fma(-x, d0c, c1_0), c0 = fma(-x, m00, c0_0);
            const double c1p = fma(-3.0 * y, y, c1);
            const double c0p = fma(y, fma(-2.0 * y, y, c1), c0);
            const RootsC R = eigen_roots_centered(c1p, c0p);
            // diagonal of M: (d0c + 2y, d1c - y, d2c - y); its distance to mu0 in FP64, rounded once
            const double o0 = fma(-2.0, y, R.m0), o12 = R.m0 + y;
            return layer_coefficients((float)(d0c - o0), (float)(d1c - o12), (float)(d2c - o12), R, t);
        } else {
            Herm3 c = herm_axpy(rho, vm, h);
            const double tr3 = (c.d0 + c.d1 + c.d2) * kTab[16];
            c.d0 -= tr3; c.d1 -= tr3; c.d2 -= tr3;
            double c2, c1, c0;
            char_poly(c, c2, c1, c0);
            const RootsC R = eigen_roots_centered(c1, c0);
            herm_offdiag_to_float(c, *Lgen);
            return layer_coefficients((float)(c.d0 - R.m0), (float)(c.d1 - R.m0), (float)(c.d2 - R.m0), R, t);
        }
    }
};

// ---- lock-step form of the FP64 side for a PAIR --------------------------------------------------------------
// eigen_roots_centered + layer_coefficients for the two events of a pair, written stage by stage over e = 0, 1 (and
// over the 4 phase arguments) so that independent instructions sit next to each other: one warp then has 2-4
// dependency chains in flight instead of one (the per-event forms above are each a single chain of ~100 dependent
// FP64 / float instructions, and with four warps per scheduler their latency, not the issue rate, set the time).
// Same operations in the same order per event: bit-identical to the per-event forms.
#define PISAB_E2 _Pragma("unroll") for (int e = 0; e < 2; ++e)
#define PISAB_E4 _Pragma("unroll") for (int e = 0; e < 4; ++e)
__device__ __forceinline__ void eigen_roots_centered2(const double (&c1)[2], const double (&c0)[2], RootsC (&R)[2]) {
    double p[2], q[2], disc[2], rs[2], b[2], inv[2], sq[2], zr[2], zi[2];
    PISAB_E2 { p[e] = -3.0 * c1[e]; p[e] = p[e] > kTab[20] ? p[e] : kTab[20]; q[e] = -13.5 * c0[e]; }
    PISAB_E2 disc[e] = 27.0 * fma(0.25 * c1[e] * c1[e], p[e] - c1[e], c0[e] * fma(6.75, c0[e], q[e]));
    { // rsqrt_fast(p), sqrt_pos(disc): seeds first, then the refinements side by side
        double r0[2], r1[2], t0[2], t1[2], e0[2], e1[2], xs[2];
        PISAB_E2 { xs[e] = disc[e] > kTab[20] ? disc[e] : kTab[20]; r0[e] = rsqrt_seed(p[e]); r1[e] = rsqrt_seed(xs[e]); }
        PISAB_E2 { t0[e] = p[e] * r0[e]; t1[e] = xs[e] * r1[e]; }
        PISAB_E2 { e0[e] = fma(-t0[e], r0[e], 1.0); e1[e] = fma(-t1[e], r1[e], 1.0); }
        PISAB_E2 { rs[e] = fma(r0[e] * e0[e], fma(e0[e], kTab[19], 0.5), r0[e]); sq[e] = fma(t1[e] * e1[e], fma(e1[e], kTab[19], 0.5), t1[e]); }
    }
    PISAB_E2 { b[e] = kTab[17] * (p[e] * rs[e]); inv[e] = rs[e] * rs[e] * rs[e]; }
    PISAB_E2 { zr[e] = q[e] * inv[e]; zi[e] = sq[e] * inv[e]; }
    float cf[2], sf[2];
    { // unit_cube_root_seed, stage-wise
        float x[2], y[2], ax[2], mx[2], mn[2], t[2], u[2], a[2], th[2], v[2];
        PISAB_E2 { x[e] = (float)zr[e]; y[e] = (float)zi[e]; ax[e] = fabsf(x[e]); mx[e] = fmaxf(ax[e], y[e]); mn[e] = fminf(ax[e], y[e]); }
#ifdef PISAB_HOST_EMU
        PISAB_E2 t[e] = mn[e] / mx[e];
#else
        PISAB_E2 t[e] = __fdividef(mn[e], mx[e]);
#endif
        PISAB_E2 { u[e] = t[e] * t[e]; a[e] = fmaf(u[e], 0.00782548263669014f, -0.03689862787723541f); }
        PISAB_E2 a[e] = fmaf(u[e], a[e], 0.08374155312776566f);
        PISAB_E2 a[e] = fmaf(u[e], a[e], -0.13480405509471893f);
        PISAB_E2 a[e] = fmaf(u[e], a[e], 0.19879871606826782f);
        PISAB_E2 a[e] = fmaf(u[e], a[e], -0.3332637548446655f);
        PISAB_E2 a[e] = fmaf(u[e], a[e], 0.9999993443489075f) * t[e];
        PISAB_E2 { a[e] = y[e] > ax[e] ? 1.57079632679489662f - a[e] : a[e]; a[e] = x[e] < 0.0f ? 3.14159265358979324f - a[e] : a[e]; }
        PISAB_E2 { th[e] = a[e] * (1.0f / 3.0f); v[e] = th[e] * th[e]; }
        PISAB_E2 {
            sf[e] = th[e] * fmaf(v[e], fmaf(v[e], fmaf(v[e], -0.00019222621631342918f, 0.008328950963914394f), -0.1666656732559204f), 0.9999999403953552f);
            cf[e] = fmaf(v[e], fmaf(v[e], fmaf(v[e], -0.0013333901297301054f, 0.041627395898103714f), -0.4999910891056061f), 0.9999997019767761f);
        }
    }
    double ct[2], st[2];
    { // unit_cube_root_mp correction
        double c[2], s2[2], m[2], rho[2], third[2], u2r[2], u2i[2], u3r[2], u3i[2], d[2];
        PISAB_E2 { c[e] = (double)cf[e]; s2[e] = (double)sf[e]; }
        PISAB_E2 { m[e] = fma(c[e], c[e], fma(s2[e], s2[e], -1.0)); u2r[e] = fma(c[e], c[e], -s2[e] * s2[e]); u2i[e] = 2.0 * c[e] * s2[e]; }
        PISAB_E2 { rho[e] = fma(m[e], -0.5, 1.0); third[e] = fma(m[e], -0.5, kTab[16]); u3r[e] = fma(u2r[e], c[e], -u2i[e] * s2[e]); u3i[e] = fma(u2r[e], s2[e], u2i[e] * c[e]); }
        PISAB_E2 d[e] = fma(zi[e], u3r[e], -zr[e] * u3i[e]) * third[e];
        PISAB_E2 { ct[e] = fma(-s2[e], d[e], c[e]) * rho[e]; st[e] = fma(c[e], d[e], s2[e]) * rho[e]; }
    }
    const double kh = 0.5, ks = kTab[15];
    PISAB_E2 { R[e].m0 = b[e] * (-kh * ct[e] - ks * st[e]); R[e].m1 = b[e] * (-kh * ct[e] + ks * st[e]); R[e].m2 = b[e] * ct[e]; }
}

// expm1i_neg for the four phase arguments of a pair-layer at once
__device__ __forceinline__ void expm1i_neg4(const double (&delta)[4], CplxF (&E)[4]) {
    double shifted[4], kd[4];
    int k[4];
    float r[4], z[4], ps[4], pc[4], sn[4], cm[4];
    PISAB_E4 { shifted[e] = fma(delta[e], kTab[0], kTab[21]); k[e] = __double2loint(shifted[e]); kd[e] = shifted[e] - kTab[21]; }
    PISAB_E4 { r[e] = (float)fma(-kd[e], kTab[1], delta[e]); z[e] = r[e] * r[e]; }
    PISAB_E4 { ps[e] = fmaf(z[e], -1.9515295891e-4f, 8.3321608736e-3f); pc[e] = fmaf(z[e], 2.443315711809948e-5f, -1.388731625493765e-3f); }
    PISAB_E4 { ps[e] = fmaf(z[e], ps[e], -1.6666654611e-1f); pc[e] = fmaf(z[e], pc[e], 4.166664568298827e-2f); }
    PISAB_E4 { sn[e] = fmaf(r[e] * z[e], ps[e], r[e]); pc[e] = fmaf(z[e], pc[e], -0.5f); }
    PISAB_E4 cm[e] = z[e] * pc[e];
    PISAB_E4 {
        const bool odd = k[e] & 1;
        float c1 = odd ? sn[e] : cm[e];
        c1 = ((k[e] + 1) & 2) ? -c1 : c1;
        const float off = (k[e] & 3) == 0 ? 0.0f : ((k[e] & 3) == 2 ? -2.0f : -1.0f);
        float sd = odd ? 1.0f + cm[e] : sn[e];
        sd = (k[e] & 2) ? -sd : sd;
        E[e] = CplxF{c1 + off, -sd};
    }
}

// layer_coefficients of both events from their roots (t[e] = 2 * 2.534 * length of event e)
__device__ __forceinline__ void layer_coefficients2(const float (&ea)[2][3], const RootsC (&R)[2], const double (&t)[2],
                                                    LayerCoefF (&K)[2]) {
    double g10d[2], g20d[2], g21d[2], delta[4];
    PISAB_E2 { g10d[e] = R[e].m1 - R[e].m0; g20d[e] = R[e].m2 - R[e].m0; g21d[e] = R[e].m2 - R[e].m1; }
    PISAB_E2 { delta[2 * e] = g10d[e] * t[e]; delta[2 * e + 1] = g20d[e] * t[e]; }
    CplxF E[4];
    expm1i_neg4(delta, E);
    float g10[2], g20[2], g21[2], r10[2], r20[2], r21[2];
    PISAB_E2 { g10[e] = (float)g10d[e]; g20[e] = (float)g20d[e]; g21[e] = (float)g21d[e]; }
    PISAB_E2 { r10[e] = rcp_f32(fmaxf(g10[e], 1e-30f)); r20[e] = rcp_f32(g20[e]); r21[e] = rcp_f32(fmaxf(g21[e], 1e-30f)); }
    PISAB_E2 {
        K[e].n1 = CplxF{E[2 * e].re * r10[e], E[2 * e].im * r10[e]};
        const CplxF f02{E[2 * e + 1].re * r20[e], E[2 * e + 1].im * r20[e]};
        K[e].n2 = CplxF{(f02.re - K[e].n1.re) * r21[e], (f02.im - K[e].n1.im) * r21[e]};
        K[e].g10 = g10[e]; K[e].g20 = g20[e]; K[e].ea0 = ea[e][0]; K[e].ea1 = ea[e][1]; K[e].ea2 = ea[e][2];
    }
}

// One event per thread
template <bool STD>
struct H0MP {
    H0Core<STD> core;
    LayerMatF base; // STD: float image of H0c's off-diagonal part and its invariants

    __device__ __forceinline__ void init(const Herm3 &h0) {
        const Herm3 c = core.init(h0);
        if (STD) {
            herm_offdiag_to_float(c, base);
            layer_products(base);
        }
    }
    __device__ __forceinline__ void layer(double rho, const Herm3 &vm, double t, Mat3F T) const {
        if (STD) {
            assemble_matrix<float>(base, core.layer(rho, vm, t, nullptr), T);
        } else {
            LayerMatF L;
            const LayerCoefF k = core.layer(rho, vm, t, &L);
            layer_products(L);
            assemble_matrix<float>(L, k, T);
        }
    }
};

__device__ __forceinline__ void pack_offdiag(const LayerMatF &a, const LayerMatF &b, LayerMatT<f2> &L) {
    L.r01 = f2{a.r01, b.r01}; L.i01 = f2{a.i01, b.i01}; L.r02 = f2{a.r02, b.r02};
    L.i02 = f2{a.i02, b.i02}; L.r12 = f2{a.r12, b.r12}; L.i12 = f2{a.i12, b.i12};
}

// Two events per thread (a pair that crosses the same Earth shells): FP64 cores per event, float part lane-packed.
// The packed per-pair invariants (15 float pairs, standard matter) live in a per-thread column of shared memory
// (`col`, [15][pitch] float2, conflict-free 64-bit accesses): in registers they left too few for the compiler to
// interleave the two events' FP64 eigenvalue chains, and a warp's time was the SERIAL latency of both chains
// (ncu, round 2: issue 61 %, `wait` stalls 33 %).
template <bool STD>
struct H0MP2 {
    H0Core<STD> core[2];
    float2 *col; // &s_base[0][threadIdx.x]
    int pitch;
    static constexpr int kSlots = 15;

    __device__ __forceinline__ void put(int k, f2 v) { col[k * pitch] = make_float2(v.x, v.y); }
    __device__ __forceinline__ f2 get(int k) const {
        const float2 v = col[k * pitch];
        return f2{v.x, v.y};
    }
    __device__ __forceinline__ void store(const LayerMatT<f2> &L) {
        put(0, L.r01); put(1, L.i01); put(2, L.r02); put(3, L.i02); put(4, L.r12); put(5, L.i12);
        put(6, L.s0); put(7, L.s1); put(8, L.s2);
        put(9, L.c01.re); put(10, L.c01.im); put(11, L.c02.re); put(12, L.c02.im); put(13, L.c12.re); put(14, L.c12.im);
    }
    __device__ __forceinline__ LayerMatT<f2> load() const {
        LayerMatT<f2> L;
        L.r01 = get(0); L.i01 = get(1); L.r02 = get(2); L.i02 = get(3); L.r12 = get(4); L.i12 = get(5);
        L.s0 = get(6); L.s1 = get(7); L.s2 = get(8);
        L.c01 = Cplx2{get(9), get(10)}; L.c02 = Cplx2{get(11), get(12)}; L.c12 = Cplx2{get(13), get(14)};
        return L;
    }
    __device__ __forceinline__ void init(const Herm3 &ha, const Herm3 &hb) {
        const Herm3 ca = core[0].init(ha), cb = core[1].init(hb);
        if (STD) {
            LayerMatF a, b;
            herm_offdiag_to_float(ca, a);
            herm_offdiag_to_float(cb, b);
            LayerMatT<f2> base;
            pack_offdiag(a, b, base);
            layer_products(base);
            store(base);
        }
    }
    // same shell for both events (same density), their own lengths
    __device__ __forceinline__ void layer(double rho, const Herm3 &vm, double ta, double tb, Cplx2 (*T)[3]) const {
        if (STD) {
#ifdef PISAB_PAIR_NO_LOCKSTEP
            const LayerCoefF ka = core[0].layer(rho, vm, ta, nullptr), kb = core[1].layer(rho, vm, tb, nullptr);
            assemble_matrix<f2>(load(), pack_coef(ka, kb), T);
#else
            // H0Core<true>::layer for both events in lock-step (see eigen_roots_centered2)
            const double x = rho * vm.d0, y = x * kTab[16];
            double c1p[2], c0p[2];
            PISAB_E2 {
                const double c1 = fma(-x, core[e].d0c, core[e].c1_0), c0 = fma(-x, core[e].m00, core[e].c0_0);
                c1p[e] = fma(-3.0 * y, y, c1);
                c0p[e] = fma(y, fma(-2.0 * y, y, c1), c0);
            }
            RootsC R[2];
            eigen_roots_centered2(c1p, c0p, R);
            float ea[2][3];
            PISAB_E2 {
                const double o0 = fma(-2.0, y, R[e].m0), o12 = R[e].m0 + y;
                ea[e][0] = (float)(core[e].d0c - o0); ea[e][1] = (float)(core[e].d1c - o12); ea[e][2] = (float)(core[e].d2c - o12);
            }
            const double tt[2] = {ta, tb};
            LayerCoefF K[2];
            layer_coefficients2(ea, R, tt, K);
            assemble_matrix<f2>(load(), pack_coef(K[0], K[1]), T);
#endif
        } else {
            LayerMatF a, b;
            const LayerCoefF ka = core[0].layer(rho, vm, ta, &a), kb = core[1].layer(rho, vm, tb, &b);
            LayerMatT<f2> L;
            pack_offdiag(a, b, L);
            layer_products(L);
            assemble_matrix<f2>(L, pack_coef(ka, kb), T);
        }
    }
};

__device__ __forceinline__ Herm3F herm_to_float(const Herm3 &h) {
    return Herm3F{(float)h.d0, (float)h.d1, (float)h.d2, (float)h.r01, (float)h.i01,
                  (float)h.r02, (float)h.i02, (float)h.r12, (float)h.i12};
}

// Vacuum columns 1 + z2 P2 + z3 P3 (see vacuum_columns) with z_k = exp(-i phi_k) - 1 (expm1i_neg); R = float: one
// event, z2 / z3 scalar; R = f2: a pair, the projector entries splatted to both lanes.
template <int NC, typename R, typename PROP>
__device__ __forceinline__ void vacuum_columns_t(const OscTable &o, CplxT<R> z2, CplxT<R> z3, PROP &P) {
    const Herm3F A = herm_to_float(o.pr2), B = herm_to_float(o.pr3);
    const R z2r = z2.re, z2i = z2.im, z3r = z3.re, z3i = z3.im;
#define PISAB_VAC_DIAG(C, DA, DB)                                                                            \
    P.set_right(C, C, CplxT<R>{t_fma(z2r, t_splat<R>(DA), t_fma(z3r, t_splat<R>(DB), t_splat<R>(1.0f))),       \
                               t_fma(z2i, t_splat<R>(DA), t_mul(z3i, t_splat<R>(DB)))});
#define PISAB_VAC_OFF(I, J, AR, AI, BR, BI)                                                                              \
    {                                                                                                                    \
        const R xr = t_fma(z2r, t_splat<R>(AR), t_mul(z3r, t_splat<R>(BR))), xi = t_fma(z2i, t_splat<R>(AR), t_mul(z3i, t_splat<R>(BR))); \
        const R yr = t_fma(z2r, t_splat<R>(AI), t_mul(z3r, t_splat<R>(BI))), yi = t_fma(z2i, t_splat<R>(AI), t_mul(z3i, t_splat<R>(BI))); \
        if (J < NC) P.set_right(J, I, CplxT<R>{t_sub(xr, yi), t_add(xi, yr)});                                           \
        if (I < NC) P.set_right(I, J, CplxT<R>{t_add(xr, yi), t_sub(xi, yr)});                                           \
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
// ts carries the nu / nubar sign: vacuum_columns uses exp(+i hdm ts) - 1 = expm1i_neg(-hdm ts)
template <int NC, typename PROP>
__device__ __forceinline__ void vacuum_columns_mp(const OscTable &o, double ts, PROP &P) {
    vacuum_columns_t<NC, float>(o, expm1i_neg(-o.hdm21 * ts), expm1i_neg(-o.hdm31 * ts), P);
}
template <int NC, typename PROP>
__device__ __forceinline__ void vacuum_columns_mp2(const OscTable &o, double tsa, double tsb, PROP &P) {
    const CplxF a2 = expm1i_neg(-o.hdm21 * tsa), a3 = expm1i_neg(-o.hdm31 * tsa);
    const CplxF b2 = expm1i_neg(-o.hdm21 * tsb), b3 = expm1i_neg(-o.hdm31 * tsb);
    vacuum_columns_t<NC, f2>(o, Cplx2{f2{a2.re, b2.re}, f2{a2.im, b2.im}}, Cplx2{f2{a3.re, b3.re}, f2{a3.im, b3.im}}, P);
}

} // namespace pisab
