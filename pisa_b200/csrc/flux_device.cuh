// flux_device.cuh -- the per-event, per-template part of flux.barr_simple, shared by the stand-alone flux kernels
// (flux.cu) and the fused template kernels (prob3.cu), which evaluate it in registers so that `nu_flux` is never
// written to or re-read from HBM in a fit that floats the flux systematics.
//
// Reference: pisa/stages/flux/barr_simple.py:139-197 (apply_sys_kernel) with the event-only transcendental terms
// t0..t3 precomputed once per container by flux_barr_terms_kernel (flux.cu).  Every operation is spelled with explicit
// rounding intrinsics: the function is inlined into kernels of very different shape and must give the same bits in
// all of them (the fused and the staged form of a template are compared bit for bit).
#pragma once
#include "common.cuh"

namespace pisab {

// the five systematic parameters of flux.barr_simple (barr_simple.py:41-52)
struct BarrSys {
    double nue_numu_ratio, nu_nubar_ratio, delta_index, uphor, nubar_sys;
};

__device__ __forceinline__ double rcp_nr(double x) { // 1/x to ~1 ulp (MUFU seed + third-order step)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    return fma(__dmul_rn(r, e), __dadd_rn(1.0, e), r);
}
// barr_simple.py:107-136 with sum_constant = True
__device__ __forceinline__ void ratio_scale_fast(double scale, double in1, double in2, double &o0, double &o1) {
    if (in1 == 0. && in2 == 0.) { o0 = 0.; o1 = 0.; return; }
    const double sr = __dmul_rn(scale, __dmul_rn(in1, rcp_nr(in2)));
    const double nw = __dmul_rn(__dadd_rn(in1, in2), rcp_nr(__dadd_rn(1., sr)));
    o0 = __dmul_rn(sr, nw);
    o1 = nw;
}

// nu_flux (e, mu) of one event from its cached terms t = (ln(E/E0), ModFlux_e, ModFlux_mu, uphor shape) and the
// nominal fluxes; ratio_scale is homogeneous of degree one, so the spectral-index factor is applied last
__device__ __forceinline__ void barr_apply_values(const BarrSys &S, double t0, double t1, double t2, double t3,
                                                  double nu0, double nu1, double nb0, double nb1, int nubar,
                                                  double &o0, double &o1) {
    double a0, a1, b0, b1, e_nu, e_nb, m_nu, m_nb;
    ratio_scale_fast(S.nue_numu_ratio, nu0, nu1, a0, a1);
    ratio_scale_fast(S.nue_numu_ratio, nb0, nb1, b0, b1);
    ratio_scale_fast(S.nu_nubar_ratio, a0, b0, e_nu, e_nb);
    ratio_scale_fast(S.nu_nubar_ratio, a1, b1, m_nu, m_nb);
    o0 = nubar < 0 ? e_nb : e_nu;
    o1 = nubar < 0 ? m_nb : m_nu;
    const double half = __dmul_rn(0.5, S.nubar_sys);
    const double h0 = __dmul_rn(half, t1), h1 = __dmul_rn(half, t2);
    const double f0 = nubar < 0 ? rcp_nr(__dadd_rn(1., h0)) : __dadd_rn(1., h0);
    const double f1 = nubar < 0 ? rcp_nr(__dadd_rn(1., h1)) : __dadd_rn(1., h1);
    o0 = __dmul_rn(o0, fmax(0., f0));
    o1 = __dmul_rn(o1, fmax(0., f1));
    o0 = __dmul_rn(o0, fma(__dmul_rn(-0.3, S.uphor), t3, 1.0));
    const double idx_scale = exp(__dmul_rn(S.delta_index, t0)); // (E/E0)^delta_index
    o0 = __dmul_rn(o0, idx_scale);
    o1 = __dmul_rn(o1, idx_scale);
}

// event i of a container: terms [n][4] doubles (32-byte aligned), nominal fluxes [n][2] of the storage type
template <typename IO>
__device__ __forceinline__ void barr_apply_event(const BarrSys &S, const double *__restrict__ terms,
                                                 const IO *__restrict__ nu_nom, const IO *__restrict__ nubar_nom,
                                                 int nubar, int64_t i, double &o0, double &o1) {
    const double2 ta = __ldg(reinterpret_cast<const double2 *>(terms) + 2 * i);
    const double2 tb = __ldg(reinterpret_cast<const double2 *>(terms) + 2 * i + 1);
    double nu0, nu1, nb0, nb1;
    if constexpr (sizeof(IO) == 8) {
        const double2 a = __ldg(reinterpret_cast<const double2 *>(nu_nom) + i);
        const double2 b = __ldg(reinterpret_cast<const double2 *>(nubar_nom) + i);
        nu0 = a.x; nu1 = a.y; nb0 = b.x; nb1 = b.y;
    } else {
        const float2 a = __ldg(reinterpret_cast<const float2 *>(nu_nom) + i);
        const float2 b = __ldg(reinterpret_cast<const float2 *>(nubar_nom) + i);
        nu0 = (double)a.x; nu1 = (double)a.y; nb0 = (double)b.x; nb1 = (double)b.y;
    }
    barr_apply_values(S, ta.x, ta.y, tb.x, tb.y, nu0, nu1, nb0, nb1, nubar, o0, o1);
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

} // namespace pisab
