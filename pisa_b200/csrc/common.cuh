// common.cuh -- error handling, launch bookkeeping and device tables shared by the kernels.
#pragma once
#ifndef PISAB_HOST_EMU // (defined only by tests/hostemu, which pre-includes its own shim to let g++ parse the device math)
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include <stdio.h>

#include "../../include/pisa_b200.h"

namespace pisab {

// ---- host-side error text (thread local) -------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define PISAB_CUDA_CHECK(expr)                                   \
    do {                                                         \
        cudaError_t _e = (expr);                                 \
        if (_e != cudaSuccess) return ::pisab::cuda_fail(_e, #expr); \
    } while (0)

// ---- launch bookkeeping -----------------------------------------------------------------
// Every kernel launch goes through note_launch() so bench.py can report `gpu_launches`, and
// (when profiling is on) through LaunchTimer so the dominant kernel is timed with CUDA
// events on the stream it was launched on.
void note_launch(int n = 1);
struct LaunchTimer {
    cudaStream_t stream;
    bool active;
    explicit LaunchTimer(cudaStream_t s);
    ~LaunchTimer();
};

// true when the *_f32 propagation entry points use the mixed-precision FP32 arithmetic (prob3_mp.cuh)
bool f32_math_mixed();

// number of SMs of the current device (cached); 0 on failure
int sm_count();

// ---- device-side parameter tables (passed by value as kernel arguments) -------------------
// Hermitian 3x3 in packed form: diagonal (real) + upper triangle (re, im).
struct Herm3 {
    double d0, d1, d2;
    double r01, i01, r02, i02, r12, i12;
};
struct Herm3F { // float image (FP32 mode)
    float d0, d1, d2;
    float r01, i01, r02, i02, r12, i12;
};

// Per-launch constants of the propagation (built on the host from pisab_osc_consts_t).
// The Hamiltonian of one layer is, for neutrinos,
//     H = hv * (1/E) + rho * vm + lr          (eV^2/GeV)
// with hv = 0.5 * U diag(0, dm21, dm31) U^dagger, vm = 0.5*1.52588e-4 * mat_pot,
// lr = 1e9 * lri_pot.  For antineutrinos the reference uses conj(U), -a*conj(V), -lri
// (numba_osc_kernels.py:208-217,435-440,650-653), i.e. H_bar = conj(hv/E - rho*vm - lr);
// since |conj(z)| = |z| and conj(exp(-iXt)) = exp(+i conj(X) t), the probabilities are those
// of H' = -hv/E + rho*vm + lr propagated with the same code, so only `hv` changes sign.
struct OscTable {
    Herm3 hv[2]; // [0] = nu, [1] = nubar (sign flipped)
    Herm3 vm;
    Herm3 lr;
    // Vacuum shortcut (shells with rho == 0, i.e. the atmosphere, when lri_pot == 0): there the
    // eigen-decomposition is the PMNS matrix itself, exp(-iHt) = 1 + sum_k (e^{-i phi_k} - 1) P_k
    // with the projectors P_k = u_k u_k^dagger (k = 2, 3) and phi_k = hdm[k] * t / E.
    Herm3 pr2, pr3;
    double hdm21, hdm31; // 0.5 * dm21, 0.5 * dm31
    double vac_ok;       // 1.0 when the shortcut is valid (lri_pot == 0), else 0.0
    double std_matter;   // 1.0 when vm = diag(a, 0, 0) (no NSI): layers only move H[0][0]
    // (keep this struct at 464 bytes: the FP64 scan kernel stages it in STATIC shared memory and fits two blocks per SM
    // with 48 bytes to spare -- 80 bytes of float projector copies here cost it half its occupancy, -40 %)
};

static_assert(sizeof(OscTable) == 464, "OscTable feeds the static shared memory of the scan kernel: see the note above");

// Neutrino decay (decay_flag == 1): 0.5 * U G U^dagger as a full complex 3x3 (row-major, re/im), [0] with
// G = mat_decay (neutrinos), [1] with G = -conj(mat_decay) (antineutrinos through the neutrino code, see tables.cu).
// Passed to the decay kernels only (prob3_decay.cuh); the standard kernels never see it.
struct DecayTable {
    double hd[2][3][3][2];
};

struct EarthTable {
    int n_radii;
    int idx_first_inner; // first shell with radius < r_detector (layers.py:91)
    double r_det;        // r_detector
    double rd2;          // r_detector * r_detector (rounded once, as the reference does)
    double rj2[PISAB_MAX_RADII];   // radii**2
    double rho[PISAB_MAX_RADII];   // electron-weighted densities
    double limit[PISAB_MAX_RADII]; // coszen_limit
};

int build_osc_table(const pisab_osc_consts_t *c, OscTable *out);
int build_earth_table(const pisab_earth_t *e, EarthTable *out);
int build_decay_table(const pisab_osc_consts_t *c, DecayTable *out);

} // namespace pisab
