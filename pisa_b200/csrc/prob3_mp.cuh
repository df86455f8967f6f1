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

struct CplxF {
    float re, im;
};
typedef CplxF Mat3F[3][3];

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

// The float image of one layer's trace-free Hamiltonian M and the invariants its square needs:
//   s_i  = sum_{j != i} |h_ij|^2                      (diagonal of M^2 minus m_i^2)
//   c_ij = h_ik h_kj, k the third index               (off-diagonal of M^2 minus (m_i + m_j) h_ij)
struct LayerMatF {
    float r01, i01, r02, i02, r12, i12;
    float s0, s1, s2;
    CplxF c01, c02, c12;
};
__device__ __forceinline__ void layer_products(LayerMatF &L) {
    const float n01 = fmaf(L.r01, L.r01, L.i01 * L.i01);
    const float n02 = fmaf(L.r02, L.r02, L.i02 * L.i02);
    const float n12 = fmaf(L.r12, L.r12, L.i12 * L.i12);
    L.s0 = n01 + n02;
    L.s1 = n01 + n12;
    L.s2 = n02 + n12;
    L.c01 = CplxF{fmaf(L.r02, L.r12, L.i02 * L.i12), fmaf(L.i02, L.r12, -L.r02 * L.i12)};  // h02 conj(h12)
    L.c02 = CplxF{fmaf(L.r01, L.r12, -L.i01 * L.i12), fmaf(L.r01, L.i12, L.i01 * L.r12)};  // h01 h12
    L.c12 = CplxF{fmaf(L.r01, L.r02, L.i01 * L.i02), fmaf(L.r01, L.i02, -L.i01 * L.r02)};  // conj(h01) h02
}

// T = 1 + n1 (M - mu0) + n2 (M - mu0)(M - mu1) = exp(-i M t) up to the global phase exp(+i mu0 t).
//   (M - mu0)(M - mu1)_ii = (m_i - mu0)(m_i - mu1) + s_i
//   (M - mu0)(M - mu1)_ij = h_ij (mu2 - m_k) + c_ij          (trace-free: m_i + m_j = -m_k, mu0 + mu1 = -mu2)
// The diagonal enters only through ea_i = m_i - mu0, which the caller forms in FP64 and rounds once (L.m* is not
// read here); m_i - mu1 = ea_i - g10 and mu2 - m_k = g20 - ea_k, so besides them only the three gaps are converted.
__device__ __forceinline__ void assemble_transition_mp(const LayerMatF &L, float ea0, float ea1, float ea2,
                                                       const RootsC &R, double t, Mat3F T) {
    const double g10d = R.m1 - R.m0, g20d = R.m2 - R.m0, g21d = R.m2 - R.m1;
    const CplxF e01 = expm1i_neg(g10d * t), e02 = expm1i_neg(g20d * t);
    const float g10 = (float)g10d, g20 = (float)g20d, g21 = (float)g21d;
    // (a gap that underflows in float would give 0 * inf: clamp; three equal roots cannot occur, see eigen_solve)
    const float r10 = rcp_f32(fmaxf(g10, 1e-30f)), r20 = rcp_f32(g20), r21 = rcp_f32(fmaxf(g21, 1e-30f));
    const CplxF n1{e01.re * r10, e01.im * r10};
    const CplxF f02{e02.re * r20, e02.im * r20};
    const CplxF n2{(f02.re - n1.re) * r21, (f02.im - n1.im) * r21};
    // ---- diagonal
#define PISAB_MP_DIAG(I, EA, S)                                                             \
    {                                                                                       \
        const float pp = fmaf(EA, EA - g10, S);                                             \
        T[I][I] = CplxF{fmaf(n2.re, pp, fmaf(n1.re, EA, 1.0f)), fmaf(n2.im, pp, n1.im * EA)}; \
    }
    PISAB_MP_DIAG(0, ea0, L.s0)
    PISAB_MP_DIAG(1, ea1, L.s1)
    PISAB_MP_DIAG(2, ea2, L.s2)
#undef PISAB_MP_DIAG
    // ---- off-diagonal pairs: T_ij = h z + n2 c, T_ji = conj(h) z + n2 conj(c), z = n1 + n2 (mu2 - m_k)
#define PISAB_MP_OFF(I, J, EAK, HR, HI, C)                                   \
    {                                                                        \
        const float u = g20 - EAK;                                           \
        const float zr = fmaf(n2.re, u, n1.re), zi = fmaf(n2.im, u, n1.im);  \
        const float s1 = fmaf(HR, zr, n2.re * C.re);                         \
        const float s2 = fmaf(HI, zi, n2.im * C.im);                         \
        const float s3 = fmaf(HR, zi, n2.im * C.re);                         \
        const float s4 = fmaf(HI, zr, n2.re * C.im);                         \
        T[I][J] = CplxF{s1 - s2, s3 + s4};                                   \
        T[J][I] = CplxF{s1 + s2, s3 - s4};                                   \
    }
    PISAB_MP_OFF(0, 1, ea2, L.r01, L.i01, L.c01)
    PISAB_MP_OFF(0, 2, ea1, L.r02, L.i02, L.c02)
    PISAB_MP_OFF(1, 2, ea0, L.r12, L.i12, L.c12)
#undef PISAB_MP_OFF
}

// Register-resident propagation state in float (same interface as Propagator / PropagatorSmem).
template <int NR, int NC>
struct PropagatorF {
    static constexpr bool kF32 = true;
    typedef CplxF cplx;
    CplxF L[NR][3];
    CplxF R[NC][3];

    __device__ __forceinline__ void set_right(int c, int k, CplxF v) { R[c][k] = v; }
    __device__ __forceinline__ void init_right(const Mat3F T) {
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int k = 0; k < 3; ++k) R[c][k] = T[k][c];
    }
    __device__ __forceinline__ void init_left(const Mat3F T, int flav) {
        if (NR == 3) {
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) L[r][c] = T[r][c];
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                CplxF v = T[0][c];
                if (flav == 1) v = T[1][c];
                if (flav == 2) v = T[2][c];
                L[0][c] = v;
            }
        }
    }
    __device__ __forceinline__ void mul_right(const Mat3F T) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const CplxF r0 = R[c][0], r1 = R[c][1], r2 = R[c][2];
#pragma unroll
            for (int k = 0; k < 3; ++k) R[c][k] = cfmaf(T[k][2], r2, cfmaf(T[k][1], r1, cmulf(T[k][0], r0)));
        }
    }
    __device__ __forceinline__ void mul_left(const Mat3F T) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const CplxF l0 = L[r][0], l1 = L[r][1], l2 = L[r][2];
#pragma unroll
            for (int c = 0; c < 3; ++c) L[r][c] = cfmaf(l2, T[2][c], cfmaf(l1, T[1][c], cmulf(l0, T[0][c])));
        }
    }
    __device__ __forceinline__ double prob(int r, int c) const {
        const CplxF acc = cfmaf(L[r][2], R[c][2], cfmaf(L[r][1], R[c][1], cmulf(L[r][0], R[c][0])));
        return (double)fmaf(acc.re, acc.re, acc.im * acc.im);
    }
};

// The same state in a per-thread column of shared memory, one (re, im) float2 per 64-bit access, [(v*3+k)][thread]:
// the in-place vector updates of a loop-carried register state cost ~190 register moves per event (ncu, round 2),
// the shared-memory form ~70 conflict-free LDS.64 / STS.64.  Vectors 0..NC-1 = columns of R, NC.. = rows of L.
template <int NR, int NC>
struct PropagatorSmemF {
    static constexpr bool kF32 = true;
    typedef CplxF cplx;
    float2 *col; // &state[0][threadIdx.x]
    int pitch;   // block size
    static constexpr int kFloat2s = (NR + NC) * 3;

    __device__ __forceinline__ CplxF ld(int v, int k) const {
        const float2 z = col[(v * 3 + k) * pitch];
        return CplxF{z.x, z.y};
    }
    __device__ __forceinline__ void st(int v, int k, CplxF z) { col[(v * 3 + k) * pitch] = make_float2(z.re, z.im); }
    __device__ __forceinline__ void set_right(int c, int k, CplxF v) { st(c, k, v); }
    __device__ __forceinline__ void init_right(const Mat3F T) {
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int k = 0; k < 3; ++k) st(c, k, T[k][c]);
    }
    __device__ __forceinline__ void init_left(const Mat3F T, int flav) {
        if (NR == 3) {
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) st(NC + r, c, T[r][c]);
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                CplxF v = T[0][c];
                if (flav == 1) v = T[1][c];
                if (flav == 2) v = T[2][c];
                st(NC, c, v);
            }
        }
    }
    __device__ __forceinline__ void mul_right(const Mat3F T) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const CplxF r0 = ld(c, 0), r1 = ld(c, 1), r2 = ld(c, 2);
#pragma unroll
            for (int k = 0; k < 3; ++k) st(c, k, cfmaf(T[k][2], r2, cfmaf(T[k][1], r1, cmulf(T[k][0], r0))));
        }
    }
    __device__ __forceinline__ void mul_left(const Mat3F T) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const CplxF l0 = ld(NC + r, 0), l1 = ld(NC + r, 1), l2 = ld(NC + r, 2);
#pragma unroll
            for (int c = 0; c < 3; ++c) st(NC + r, c, cfmaf(l2, T[2][c], cfmaf(l1, T[1][c], cmulf(l0, T[0][c]))));
        }
    }
    __device__ __forceinline__ double prob(int r, int c) const {
        const CplxF acc = cfmaf(ld(NC + r, 2), ld(c, 2), cfmaf(ld(NC + r, 1), ld(c, 1), cmulf(ld(NC + r, 0), ld(c, 0))));
        return (double)fmaf(acc.re, acc.re, acc.im * acc.im);
    }
};

// Per-event Hamiltonian provider of the FP32 mode.
//   STD (standard matter potential, vm = diag(a, 0, 0)): per event the trace-free H0c = hv/E + lr - tr/3 is kept in
//   float together with s_i, c_ij (which do not depend on the density) and, in FP64, the invariants that make the
//   cubic of H0c + x e00 - x/3 linear in x = rho * a:
//       c1(x) = c1_0 - x d0c - x^2/3 ,   c0(x) = c0_0 - x m00 + (x/3) c1(x) + ... (see layer_poly)
//   General (NSI): the layer Hamiltonian is formed and centred in FP64 per layer, then rounded.
template <bool STD>
struct H0MP {
    LayerMatF base; // STD: float image of H0c and its invariants; general: scratch
    Herm3 h;        // general: H0 = hv/E + lr in FP64 (STD: unused after init)
    double c1_0, c0_0, d0c, d1c, d2c, m00; // STD only

    __device__ __forceinline__ void init(const Herm3 &h0) {
        if (STD) {
            const double tr3 = (h0.d0 + h0.d1 + h0.d2) * kTab[16];
            Herm3 c = h0;
            c.d0 -= tr3; c.d1 -= tr3; c.d2 -= tr3;
            double c2;
            char_poly(c, c2, c1_0, c0_0); // c2 == 0 up to rounding
            d0c = c.d0; d1c = c.d1; d2c = c.d2;
            m00 = fma(c.d1, c.d2, -fma(c.r12, c.r12, c.i12 * c.i12));
            base.r01 = (float)c.r01; base.i01 = (float)c.i01; base.r02 = (float)c.r02;
            base.i02 = (float)c.i02; base.r12 = (float)c.r12; base.i12 = (float)c.i12;
            layer_products(base);
        } else {
            h = h0;
        }
    }
    // transition matrix of a layer of density rho, t = 2 * 2.534 * length
    __device__ __forceinline__ void layer(double rho, const Herm3 &vm, double t, Mat3F T) const {
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
            assemble_transition_mp(base, (float)(d0c - o0), (float)(d1c - o12), (float)(d2c - o12), R, t, T);
        } else {
            Herm3 c = herm_axpy(rho, vm, h);
            const double tr3 = (c.d0 + c.d1 + c.d2) * kTab[16];
            c.d0 -= tr3; c.d1 -= tr3; c.d2 -= tr3;
            double c2, c1, c0;
            char_poly(c, c2, c1, c0);
            const RootsC R = eigen_roots_centered(c1, c0);
            LayerMatF L;
            L.r01 = (float)c.r01; L.i01 = (float)c.i01; L.r02 = (float)c.r02;
            L.i02 = (float)c.i02; L.r12 = (float)c.r12; L.i12 = (float)c.i12;
            layer_products(L);
            assemble_transition_mp(L, (float)(c.d0 - R.m0), (float)(c.d1 - R.m0), (float)(c.d2 - R.m0), R, t, T);
        }
    }
};

__device__ __forceinline__ Herm3F herm_to_float(const Herm3 &h) {
    return Herm3F{(float)h.d0, (float)h.d1, (float)h.d2, (float)h.r01, (float)h.i01,
                  (float)h.r02, (float)h.i02, (float)h.r12, (float)h.i12};
}

// Vacuum columns 1 + z2 P2 + z3 P3 (see vacuum_columns) with z_k = exp(-i phi_k) - 1 from expm1i_neg.
template <int NC, typename PROP>
__device__ __forceinline__ void vacuum_columns_mp(const OscTable &o, double ts, PROP &P) {
    // ts carries the nu / nubar sign: exp(-i hdm ts) with ts = -/+ t / E  (vacuum_columns)
    const CplxF z2 = expm1i_neg(-o.hdm21 * ts), z3 = expm1i_neg(-o.hdm31 * ts);
    // vacuum_columns uses (cos - 1, +sin) of (hdm * ts): exp(+i hdm ts) - 1 = expm1i_neg(-hdm ts)
    const float z2r = z2.re, z2i = z2.im, z3r = z3.re, z3i = z3.im;
    const Herm3F A = herm_to_float(o.pr2), B = herm_to_float(o.pr3);
#define PISAB_VAC_DIAG(C, DA, DB) \
    P.set_right(C, C, CplxF{fmaf(z2r, DA, fmaf(z3r, DB, 1.0f)), fmaf(z2i, DA, z3i * DB)});
#define PISAB_VAC_OFF(I, J, AR, AI, BR, BI)                                                                   \
    {                                                                                                         \
        const float xr = fmaf(z2r, AR, z3r * BR), xi = fmaf(z2i, AR, z3i * BR);                                 \
        const float yr = fmaf(z2r, AI, z3r * BI), yi = fmaf(z2i, AI, z3i * BI);                                 \
        if (J < NC) P.set_right(J, I, CplxF{xr - yi, xi + yr});                                               \
        if (I < NC) P.set_right(I, J, CplxF{xr + yi, xi - yr});                                               \
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

} // namespace pisab
