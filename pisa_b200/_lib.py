"""ctypes binding of libpisa_b200.so (the C ABI declared in include/pisa_b200.h).

There is no CPU fallback: if the CUDA library has not been built, or no CUDA device is
present when a compute entry point is called, this raises.
"""
import ctypes
import os

import numpy as np

from . import build as _build

MAX_RADII = 64
MAX_LAYERS = 120
MAX_DIMS = 4
DET_MAX_BINS = 1024
DIM_LIN, DIM_LOG, DIM_EDGES = 0, 1, 2
F32_MATH_FP64, F32_MATH_MIXED = 0, 1
CONTAINER_PAIR_ALIGNED = 1
CONTAINER_FLUX_SYS = 2

c_i32, c_i64, c_dbl, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p


class OscConsts(ctypes.Structure):
    """pisab_osc_consts_t -- the argument list of propagate_array (numba_osc_hostfuncs.py:60-70)."""
    _fields_ = [("dm", c_dbl * 9), ("mix", c_dbl * 18), ("mat_pot", c_dbl * 18),
                ("mat_decay", c_dbl * 18), ("lri_pot", c_dbl * 9), ("decay_flag", c_i64)]

    @classmethod
    def from_matrices(cls, dm, mix, mat_pot, decay_flag=-1, mat_decay=None, lri_pot=None):
        def cplx(m):
            m = np.asarray(m, dtype=np.complex128).reshape(3, 3)
            return np.stack([m.real, m.imag], axis=-1).ravel()
        c = cls()
        c.dm[:] = np.asarray(dm, dtype=np.float64).reshape(9)
        c.mix[:] = cplx(mix)
        c.mat_pot[:] = cplx(mat_pot)
        c.mat_decay[:] = cplx(np.zeros((3, 3)) if mat_decay is None else mat_decay)
        c.lri_pot[:] = np.asarray(np.zeros((3, 3)) if lri_pot is None else lri_pot, dtype=np.float64).reshape(9)
        c.decay_flag = int(decay_flag)
        return c


class Earth(ctypes.Structure):
    """pisab_earth_t -- what Layers holds after __init__/setElecFrac (layers.py:216-289,308-335)."""
    _fields_ = [("n_radii", c_i32), ("max_layers", c_i32), ("r_detector", c_dbl),
                ("radii", c_dbl * MAX_RADII), ("rho_e", c_dbl * MAX_RADII),
                ("coszen_limit", c_dbl * MAX_RADII)]

    @classmethod
    def from_arrays(cls, radii, rho_e, coszen_limit, r_detector, max_layers):
        n = len(radii)
        if n > MAX_RADII:
            raise ValueError("Earth model has %d shells, at most %d supported" % (n, MAX_RADII))
        e = cls()
        e.n_radii, e.max_layers, e.r_detector = n, int(max_layers), float(r_detector)
        e.radii[:n] = np.asarray(radii, dtype=np.float64)
        e.rho_e[:n] = np.asarray(rho_e, dtype=np.float64)
        e.coszen_limit[:n] = np.asarray(coszen_limit, dtype=np.float64)
        return e


class Binning(ctypes.Structure):
    """pisab_binning_t -- regularised output binning (hist.py:86-127)."""
    _fields_ = [("n_dims", c_i32), ("kind", c_i32 * MAX_DIMS), ("n_bins", c_i32 * MAX_DIMS),
                ("lo", c_dbl * MAX_DIMS), ("hi", c_dbl * MAX_DIMS), ("d_edges", c_vp * MAX_DIMS)]


class ContainerDesc(ctypes.Structure):
    """pisab_container_t -- one flavour container of a batched template evaluation."""
    _fields_ = [("d_energy", c_vp), ("d_coszen", c_vp), ("d_nu_flux", c_vp), ("d_weights", c_vp),
                ("d_index", c_vp), ("d_order", c_vp), ("d_weights_out", c_vp), ("n", c_i64),
                ("scale", c_dbl), ("nubar", c_i32), ("flav", c_i32), ("flags", c_i32), ("pad", c_i32),
                ("d_flux_terms", c_vp), ("d_nu_flux_nominal", c_vp), ("d_nubar_flux_nominal", c_vp),
                ("d_astro_weights", c_vp)]


class FluxSys(ctypes.Structure):
    """pisab_flux_sys_t -- the five systematic parameters of flux.barr_simple (barr_simple.py:41-52)."""
    _fields_ = [("nue_numu_ratio", c_dbl), ("nu_nubar_ratio", c_dbl), ("delta_index", c_dbl),
                ("barr_uphor_ratio", c_dbl), ("barr_nu_nubar_ratio", c_dbl)]


class FluxItem(ctypes.Structure):
    """pisab_flux_item_t -- one flavour container of a batched flux.barr_simple evaluation."""
    _fields_ = [("d_terms", c_vp), ("d_nu_flux_nominal", c_vp), ("d_nubar_flux_nominal", c_vp), ("d_nu_flux", c_vp),
                ("n", c_i64), ("nubar", c_i32), ("pad", c_i32)]


MAX_BATCH = 16

_lib = None

_SIGNATURES = {
    # name: (restype, argtypes)
    "pisab_last_error": (ctypes.c_char_p, []),
    "pisab_version": (ctypes.c_char_p, []),
    "pisab_device_info": (c_i32, [ctypes.POINTER(c_i32)] * 3),
    "pisab_layers_calc": (c_i32, [ctypes.POINTER(Earth), c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "pisab_prob3_propagate_layers": (c_i32, [ctypes.POINTER(OscConsts), c_i32, c_vp, c_vp, c_vp, c_vp, c_i64,
                                             c_i32, c_vp, c_vp]),
    "pisab_prob3_propagate_earth": (c_i32, [ctypes.POINTER(OscConsts), ctypes.POINTER(Earth), c_i32, c_vp, c_i32,
                                            c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "pisab_layer_count": (c_i32, [ctypes.POINTER(Earth), c_vp, c_i64, c_vp, c_vp]),
    "pisab_fill_probs": (c_i32, [c_vp, c_i32, c_i32, c_i64, c_vp, c_vp]),
    "pisab_apply_osc_weights": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp]),
    "pisab_hist_index": (c_i32, [ctypes.POINTER(Binning), ctypes.POINTER(c_vp), c_i64, c_vp, c_vp]),
    "pisab_hist_accumulate": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "pisab_lookup": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_vp, c_vp]),
    "pisab_reweight_hist": (c_i32, [ctypes.POINTER(OscConsts), ctypes.POINTER(Earth), c_i32, c_vp, c_i32, c_vp,
                                    c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp,
                                    c_vp, c_i64, c_vp]),
    "pisab_flux_barr_simple": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i32, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl, c_i64, c_vp,
                                       c_vp]),
    "pisab_flux_barr_terms": (c_i32, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    "pisab_flux_barr_apply": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl, c_i64, c_vp, c_vp]),
    "pisab_flux_barr_apply_batch": (c_i32, [ctypes.POINTER(FluxItem), c_i32, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl, c_vp]),
    "pisab_flux_honda_2d": (c_i32, [c_vp, c_i32, c_vp, c_i32, c_vp, c_i32, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "pisab_reweight_hist_scan": (c_i32, [ctypes.POINTER(OscConsts), c_i32, ctypes.POINTER(Earth),
                                         ctypes.POINTER(ContainerDesc), c_i32, c_i32, c_vp, c_vp, c_i64, c_vp]),
    "pisab_hist_accumulate_planned": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "pisab_hist_accumulate_sorted": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "pisab_scale_weights": (c_i32, [c_vp, c_dbl, c_i64, c_vp, c_vp]),
    "pisab_hist_transform": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "pisab_reweight_hist_chi2": (c_i32, [ctypes.POINTER(OscConsts), ctypes.POINTER(Earth),
                                         ctypes.POINTER(ContainerDesc), c_i32, c_i32, ctypes.POINTER(FluxSys), c_vp,
                                         c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "pisab_reweight_hist_batch": (c_i32, [ctypes.POINTER(OscConsts), ctypes.POINTER(Earth),
                                          ctypes.POINTER(ContainerDesc), c_i32, c_i32, ctypes.POINTER(FluxSys), c_vp,
                                          c_vp, c_i64, c_vp]),
}
_UNTYPED = {
    "pisab_hist_workspace_bytes": (c_i64, [c_i64, c_i32]),
    "pisab_reweight_batch_workspace_bytes": (c_i64, [c_i32, c_i32]),
    "pisab_hist_plan_bytes": (c_i64, [c_i64, c_i32]),
    "pisab_sort_workspace_bytes": (c_i64, [c_i64]),
    "pisab_sort_order_i32": (c_i32, [c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "pisab_hist_plan_build": (c_i32, [c_vp, c_i64, c_i32, c_vp, c_i64, c_vp]),
    "pisab_exchange_create": (c_i32, [c_i32, c_i32, c_i64, ctypes.POINTER(c_vp), c_vp]),
    "pisab_exchange_connect": (c_i32, [c_vp, c_vp]),
    "pisab_exchange_allreduce": (c_i32, [c_vp, c_vp, c_i64, c_vp]),
    "pisab_exchange_status": (c_i32, [c_vp]),
    "pisab_exchange_disconnect": (c_i32, [c_vp]),
    "pisab_exchange_destroy": (c_i32, [c_vp]),
    "pisab_sum_slots": (c_i32, [c_vp, c_i32, c_i64, c_vp, c_vp]),
    "pisab_joint_index": (c_i32, [c_vp, c_vp, c_i32, c_i64, c_vp, c_vp]),
    "pisab_mod_chi2": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_vp, c_vp]),
    "pisab_template_chi2": (c_i32, [c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "pisab_hist_scale_sum_chi2": (c_i32, [c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "pisab_template_chi2_batch": (c_i32, [c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "pisab_reweight_scan_workspace_bytes": (c_i64, [c_i32, c_i32, c_i32, c_i64]),
    "pisab_fp64_peak_probe": (c_i32, [c_i32, ctypes.POINTER(c_dbl), ctypes.POINTER(c_dbl)]),
    "pisab_launch_count": (c_i64, [c_i32]),
    "pisab_set_profiling": (c_i32, [c_i32]),
    "pisab_set_f32_math": (c_i32, [c_i32]),
    "pisab_get_f32_math": (c_i32, []),
    "pisab_last_kernel_ms": (c_dbl, []),
}
_NO_SUFFIX = ("pisab_last_error", "pisab_version", "pisab_device_info")

# every symbol include/pisa_b200.h declares
EXPORTED_SYMBOLS = sorted(
    [n for n in _NO_SUFFIX]
    + [n + s for n in _SIGNATURES if n not in _NO_SUFFIX for s in ("_f64", "_f32")]
    + list(_UNTYPED)
)


class PisabError(RuntimeError):
    """Any non-zero status of the C ABI (CUDA failure, workspace too small, ...)."""


class PisabArgError(PisabError, ValueError):
    """PISAB_ERR_ARG: the exception class the reference raises for bad modes / params (stage.py:360-379)."""


class PisabUnsupportedError(PisabError, NotImplementedError):
    """PISAB_ERR_UNSUPPORTED: a branch of the reference that is outside the hot path (e.g. a non-Hermitian matter potential) or a combination an entry point does not fuse."""


_ERR_CLASSES = {1: PisabArgError, 3: PisabUnsupportedError}


def lib_path():
    return _build.LIB


def load():
    """Load the CUDA library; raise (never fall back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(
            "pisa_b200: CUDA library %s is missing and there is no CPU fallback. "
            "Build it with `python -c 'import __graft_entry__ as g; g.build()'`." % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        names = [name] if name in _NO_SUFFIX else [name + "_f64", name + "_f32"]
        for n in names:
            fn = getattr(lib, n)
            fn.restype, fn.argtypes = res, args
    for name, (res, args) in _UNTYPED.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise _ERR_CLASSES.get(rc, PisabError)("pisa_b200 error %d: %s" % (rc, load().pisab_last_error().decode()))


def fn(name, dtype):
    """Typed entry point for a torch / numpy float dtype."""
    sfx = {"float64": "_f64", "float32": "_f32"}[str(dtype).replace("torch.", "")]
    return getattr(load(), name + sfx)
