"""Device operators of the hot path: thin, typed wrappers over the C ABI (include/pisa_b200.h).

All array arguments are CUDA ``torch`` tensors (torch is used for device memory and streams
only); parameter structs are host objects.  Every call enqueues on torch's current stream.
There is no CPU path: a non-CUDA tensor raises.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import MAX_BATCH, Binning, ContainerDesc, Earth, OscConsts  # noqa: F401  (re-exported)

_FLOATS = (torch.float64, torch.float32)


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    """torch's current stream on the current device as a cudaStream_t (the raw accessor avoids ~15 us of
    Stream-object construction per call: a template is ~100 operator calls through the Stage API)."""
    if _raw_stream is not None:
        return ctypes.c_void_p(_raw_stream(torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(t, name, dtype=None, allow_none=False):
    if t is None:
        if allow_none:
            return None
        raise ValueError("%s is required" % name)
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError("%s must be a CUDA torch tensor (pisa_b200 has no CPU path)" % name)
    if dtype is not None and t.dtype != dtype:
        raise TypeError("%s must have dtype %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t


def _ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _check_order(order, n):
    if order is not None:
        _chk(order, "order", torch.int32)
        if order.numel() != n:
            raise ValueError("order must be a permutation of all %d events" % n)


def _species(val, n, name):
    """(+1|-1 or flavour) given as python int (aux scalar of a container) or int32 tensor [n]."""
    if isinstance(val, torch.Tensor):
        _chk(val, name, torch.int32)
        if val.numel() != n:
            raise ValueError("%s must have one entry per event" % name)
        return 0, val
    return int(val), None


def set_f32_math(mode):
    """Arithmetic behind float32 event arrays: "mixed" (default; FP64 eigenvalues / phases, float32 matrices and
    state: the FP32 mode) or "fp64" (float32 storage only).  Process-wide (``pisab_set_f32_math``)."""
    code = {"mixed": _lib.F32_MATH_MIXED, "fp64": _lib.F32_MATH_FP64}.get(mode, mode)
    _lib.check(_lib.load().pisab_set_f32_math(int(code)))


def get_f32_math():
    return {_lib.F32_MATH_MIXED: "mixed", _lib.F32_MATH_FP64: "fp64"}[int(_lib.load().pisab_get_f32_math())]


def device_info():
    sms, maj, mnr = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    _lib.check(_lib.load().pisab_device_info(ctypes.byref(sms), ctypes.byref(maj), ctypes.byref(mnr)))
    return dict(sm_count=sms.value, cc=(maj.value, mnr.value))


# ------------------------------------------------------------------------------- flux -----

def flux_barr_simple(true_energy, true_coszen, nu_flux_nominal, nubar_flux_nominal, nubar, nue_numu_ratio,
                     nu_nubar_ratio, delta_index, Barr_uphor_ratio, Barr_nu_nubar_ratio, out=None):
    """``apply_sys_vectorized`` of flux.barr_simple (barr_simple.py:200-226): returns ``nu_flux`` [n, 2]."""
    _chk(true_energy, "true_energy")
    dt = true_energy.dtype
    n = true_energy.numel()
    for t, nm in ((true_coszen, "true_coszen"), (nu_flux_nominal, "nu_flux_nominal"),
                  (nubar_flux_nominal, "nubar_flux_nominal")):
        _chk(t, nm, dt)
    if nu_flux_nominal.shape != (n, 2) or nubar_flux_nominal.shape != (n, 2) or true_coszen.numel() != n:
        raise ValueError("inconsistent event array shapes")
    if out is None:
        out = torch.empty((n, 2), dtype=dt, device=true_energy.device)
    _chk(out, "out", dt)
    f = _lib.fn("pisab_flux_barr_simple", dt)
    _lib.check(f(_ptr(true_energy), _ptr(true_coszen), _ptr(nu_flux_nominal), _ptr(nubar_flux_nominal), int(nubar),
                 float(nue_numu_ratio), float(nu_nubar_ratio), float(delta_index), float(Barr_uphor_ratio),
                 float(Barr_nu_nubar_ratio), n, _ptr(out), _stream()))
    return out


def flux_barr_terms(true_energy, true_coszen, out=None):
    """The four parameter-independent per-event terms of flux.barr_simple (float64 [n, 4]); see
    ``pisab_flux_barr_terms_*``.  Computed once per container; ``flux_barr_apply`` uses them per template."""
    _chk(true_energy, "true_energy")
    _chk(true_coszen, "true_coszen", true_energy.dtype)
    n = true_energy.numel()
    if true_coszen.numel() != n:
        raise ValueError("inconsistent event array shapes")
    if out is None:
        out = torch.empty((n, 4), dtype=torch.float64, device=true_energy.device)
    _chk(out, "out", torch.float64)
    f = _lib.fn("pisab_flux_barr_terms", true_energy.dtype)
    _lib.check(f(_ptr(true_energy), _ptr(true_coszen), n, _ptr(out), _stream()))
    return out


def flux_barr_apply(terms, nu_flux_nominal, nubar_flux_nominal, nubar, nue_numu_ratio, nu_nubar_ratio, delta_index,
                    Barr_uphor_ratio, Barr_nu_nubar_ratio, out=None):
    """``nu_flux`` [n, 2] from precomputed ``flux_barr_terms`` and the nominal fluxes: the fit-loop form of
    ``flux_barr_simple`` (same result to rounding, HBM-bound)."""
    _chk(terms, "terms", torch.float64)
    _chk(nu_flux_nominal, "nu_flux_nominal")
    dt = nu_flux_nominal.dtype
    _chk(nubar_flux_nominal, "nubar_flux_nominal", dt)
    n = terms.shape[0]
    if terms.shape != (n, 4) or nu_flux_nominal.shape != (n, 2) or nubar_flux_nominal.shape != (n, 2):
        raise ValueError("inconsistent event array shapes")
    if out is None:
        out = torch.empty((n, 2), dtype=dt, device=terms.device)
    _chk(out, "out", dt)
    f = _lib.fn("pisab_flux_barr_apply", dt)
    _lib.check(f(_ptr(terms), _ptr(nu_flux_nominal), _ptr(nubar_flux_nominal), int(nubar), float(nue_numu_ratio),
                 float(nu_nubar_ratio), float(delta_index), float(Barr_uphor_ratio), float(Barr_nu_nubar_ratio), n,
                 _ptr(out), _stream()))
    return out


class FluxBatch:
    """Descriptor array for ``flux_barr_apply_batch``: the flux.barr_simple re-evaluation of up to MAX_BATCH containers
    in one launch.  items: dicts with terms, nu_flux_nominal, nubar_flux_nominal, nu_flux (output), nubar."""

    def __init__(self, items):
        if not 1 <= len(items) <= MAX_BATCH:
            raise ValueError("a batch holds 1..%d containers" % MAX_BATCH)
        self.n = len(items)
        self.desc = (_lib.FluxItem * self.n)()
        self._keep, dt = [], None
        for d, c in zip(self.desc, items):
            t = _chk(c["terms"], "terms", torch.float64)
            nu = _chk(c["nu_flux_nominal"], "nu_flux_nominal")
            dt = dt or nu.dtype
            nb, out = _chk(c["nubar_flux_nominal"], "nubar_flux_nominal", dt), _chk(c["nu_flux"], "nu_flux", dt)
            n = t.shape[0]
            if t.shape != (n, 4) or nu.shape != (n, 2) or nb.shape != (n, 2) or out.shape != (n, 2):
                raise ValueError("inconsistent event array shapes")
            d.d_terms, d.d_nu_flux_nominal, d.d_nubar_flux_nominal = t.data_ptr(), nu.data_ptr(), nb.data_ptr()
            d.d_nu_flux, d.n, d.nubar = out.data_ptr(), n, int(c["nubar"])
            self._keep.append((t, nu, nb, out))
        self.dtype = dt


def flux_barr_apply_batch(batch, nue_numu_ratio, nu_nubar_ratio, delta_index, Barr_uphor_ratio, Barr_nu_nubar_ratio):
    f = _lib.fn("pisab_flux_barr_apply_batch", batch.dtype)
    _lib.check(f(batch.desc, batch.n, float(nue_numu_ratio), float(nu_nubar_ratio), float(delta_index),
                 float(Barr_uphor_ratio), float(Barr_nu_nubar_ratio), _stream()))


def flux_honda_2d(table, true_energy, true_coszen, nu_flux_nominal=None, nubar_flux_nominal=None):
    """``calculate_2d_flux_weights`` (flux_weights.py:267-350) for the four primaries of ``table``
    (a ``pisa_b200.utils.flux_weights.HondaTable2D``): returns (nu_flux_nominal, nubar_flux_nominal), [n, 2] each."""
    _chk(true_energy, "true_energy")
    dt = true_energy.dtype
    _chk(true_coszen, "true_coszen", dt)
    n = true_energy.numel()
    if true_coszen.numel() != n:
        raise ValueError("length of energy and coszen arrays must match")
    if n and not bool(((true_coszen >= -1.0) & (true_coszen <= 1.0)).all()):
        raise ValueError("Not all coszens found between -1 and 1")      # flux_weights.py:323-324
    if nu_flux_nominal is None:
        nu_flux_nominal = torch.empty((n, 2), dtype=dt, device=true_energy.device)
    if nubar_flux_nominal is None:
        nubar_flux_nominal = torch.empty((n, 2), dtype=dt, device=true_energy.device)
    _chk(nu_flux_nominal, "nu_flux_nominal", dt)
    _chk(nubar_flux_nominal, "nubar_flux_nominal", dt)
    knots, breaks, cells = table.device_tables(true_energy.device)
    f = _lib.fn("pisab_flux_honda_2d", dt)
    _lib.check(f(_ptr(knots), knots.numel(), _ptr(breaks), breaks.numel(), _ptr(cells), int(table.enpow),
                 _ptr(true_energy), _ptr(true_coszen), n, _ptr(nu_flux_nominal), _ptr(nubar_flux_nominal), _stream()))
    return nu_flux_nominal, nubar_flux_nominal


# ----------------------------------------------------------------------------- layers -----

def layers_calc(earth, coszen):
    """``Layers.calcLayers`` (layers.py:339-363): returns (n_layers[int32], densities, distances),
    the latter two shaped [N, earth.max_layers]."""
    _chk(coszen, "coszen")
    if coszen.dtype not in _FLOATS:
        raise TypeError("coszen must be float32/float64")
    n = coszen.numel()
    den = torch.empty((n, earth.max_layers), dtype=coszen.dtype, device=coszen.device)
    dis = torch.empty_like(den)
    nl = torch.empty(n, dtype=torch.int32, device=coszen.device)
    f = _lib.fn("pisab_layers_calc", coszen.dtype)
    _lib.check(f(ctypes.byref(earth), _ptr(coszen), n, _ptr(den), _ptr(dis), _ptr(nl), _stream()))
    return nl, den, dis


# ------------------------------------------------------------------------ propagation -----

def propagate_layers(consts, nubar, energy, densities, distances, out=None):
    """``propagate_array`` (numba_osc_hostfuncs.py:60-70) with explicit layer arrays.
    Returns probability[N,3,3] with out[i,j] = P(nu_i -> nu_j)."""
    _chk(energy, "energy")
    dt = energy.dtype
    n = energy.numel()
    _chk(densities, "densities", dt)
    _chk(distances, "distances", dt)
    if densities.shape != distances.shape or densities.shape[0] != n:
        raise ValueError("densities/distances must be [N, n_layers]")
    nl = densities.shape[1]
    nb, d_nb = _species(nubar, n, "nubar")
    if out is None:
        out = torch.empty((n, 3, 3), dtype=dt, device=energy.device)
    _chk(out, "out", dt)
    f = _lib.fn("pisab_prob3_propagate_layers", dt)
    _lib.check(f(ctypes.byref(consts), nb, _ptr(d_nb), _ptr(energy), _ptr(densities), _ptr(distances), n, nl,
                 _ptr(out), _stream()))
    return out


def layer_counts(earth, coszen):
    """Number of Earth shells every event crosses (int32; ``pisab_layer_count``): the class key of ``layer_order``."""
    _chk(coszen, "coszen")
    n = coszen.numel()
    count = torch.empty(n, dtype=torch.int32, device=coszen.device)
    f = _lib.fn("pisab_layer_count", coszen.dtype)
    _lib.check(f(ctypes.byref(earth), _ptr(coszen), n, _ptr(count), _stream()))
    return count


def sort_order(keys, descending=False, key_bits=0, want_sorted=False):
    """Stable order of non-negative int32 ``keys`` (``pisab_sort_order_i32``, radix sort): int32 permutation, and
    ``keys[order]`` with ``want_sorted``.  Setup-time helper (event grouping, sorted histogram plans)."""
    _chk(keys, "keys", torch.int32)
    n = keys.numel()
    order = torch.empty(n, dtype=torch.int32, device=keys.device)
    out = torch.empty(n, dtype=torch.int32, device=keys.device) if want_sorted else None
    if n:
        nbytes = int(_lib.load().pisab_sort_workspace_bytes(n))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=keys.device)
        _lib.check(_lib.load().pisab_sort_order_i32(_ptr(keys), n, int(key_bits), int(bool(descending)), _ptr(order),
                                                    _ptr(out), _ptr(ws), nbytes, _stream()))
    return (order, out) if want_sorted else order


def _class_order(count, index=None):
    """Stable permutation (int64) grouping the events by ``count`` (descending); with ``index`` the events of a class
    are ordered by bin as well (large binnings: a warp's 32 events then share a bin, see ``warp_fixed_add``)."""
    if index is None or count.numel() == 0:
        return sort_order(count, descending=True, key_bits=8).long()
    span = int(index.max()) + 2
    key = (int(count.max()) - count) * span + torch.clamp(index + 1, min=0)
    return sort_order(key.to(torch.int32).contiguous()).long()


def pair_aligned_order(earth, coszen, index=None):
    """Setup-time helper for the two-events-per-thread FP32 kernel: (sel, dummy) with ``sel`` (int64) listing the
    events grouped by crossed shells like ``layer_order`` and every class padded to an EVEN size by repeating its last
    event, ``dummy`` (bool) marking the repeats (their weight must be set to 0 and their bin index to -1).  Events
    2k and 2k+1 of ``x[sel]`` then cross the same shells (PISAB_CONTAINER_PAIR_ALIGNED)."""
    count = layer_counts(earth, coszen)
    order = _class_order(count, index)
    _, sizes = torch.unique_consecutive(count[order], return_counts=True)
    ends = torch.cumsum(sizes, 0) - 1
    extra = ends[sizes % 2 == 1]                       # position (in sorted order) of the event to repeat
    pos = torch.cat([torch.arange(order.numel(), device=order.device), extra])
    pos = sort_order(pos.to(torch.int32), want_sorted=True)[1].long()   # repeats land right behind their originals
    dummy = torch.zeros(pos.numel(), dtype=torch.bool, device=order.device)
    if pos.numel() > 1:
        dummy[1:] = pos[1:] == pos[:-1]
    return order[pos], dummy


def layer_order(earth, coszen, index=None):
    """Setup-time helper: permutation (int32) listing the events grouped by the number of Earth
    shells they cross, deepest first, stable within a group.  Passing it as ``order`` to
    propagate_earth / reweight_hist makes every warp walk the same number of layers; it depends on
    ``coszen`` only (the reference computes its layer arrays once in setup_function as well).
    ``index`` (the events' bin index, optional): secondary key, events of a group ordered by bin."""
    _chk(coszen, "coszen")
    n = coszen.numel()
    count = torch.empty(n, dtype=torch.int32, device=coszen.device)
    f = _lib.fn("pisab_layer_count", coszen.dtype)
    _lib.check(f(ctypes.byref(earth), _ptr(coszen), n, _ptr(count), _stream()))
    # (stable sort => the permutation, hence every sum, is reproducible)
    order = _class_order(count, index)
    return order.to(torch.int32).contiguous()


def propagate_earth(consts, earth, nubar, energy, coszen, flav=None, probability=None, prob_e=None,
                    prob_mu=None, want_probability=True, order=None):
    """Layers evaluated in-kernel from ``coszen`` (prob3.py:406-409 + :581-605 fused).

    Returns (probability | None, prob_e | None, prob_mu | None).  ``flav`` (0/1/2, python int or
    int32 tensor) selects the final flavour for prob_e / prob_mu (fill_probs)."""
    _chk(energy, "energy")
    dt = energy.dtype
    _chk(coszen, "coszen", dt)
    n = energy.numel()
    if coszen.numel() != n:
        raise ValueError("energy and coszen must have the same length")
    nb, d_nb = _species(nubar, n, "nubar")
    if want_probability and probability is None:
        probability = torch.empty((n, 3, 3), dtype=dt, device=energy.device)
    fl, d_fl = 0, None
    if flav is not None:
        fl, d_fl = _species(flav, n, "flav")
        if prob_e is None:
            prob_e = torch.empty(n, dtype=dt, device=energy.device)
        if prob_mu is None:
            prob_mu = torch.empty(n, dtype=dt, device=energy.device)
    for t, nm in ((probability, "probability"), (prob_e, "prob_e"), (prob_mu, "prob_mu")):
        _chk(t, nm, dt, allow_none=True)
    _check_order(order, n)
    f = _lib.fn("pisab_prob3_propagate_earth", dt)
    _lib.check(f(ctypes.byref(consts), ctypes.byref(earth), nb, _ptr(d_nb), fl, _ptr(d_fl), _ptr(energy),
                 _ptr(coszen), _ptr(order), n, _ptr(probability), _ptr(prob_e), _ptr(prob_mu), _stream()))
    return probability, prob_e, prob_mu


def fill_probs(probability, initial_flav, flav, out=None):
    """``fill_probs`` (numba_osc_hostfuncs.py:206-221)."""
    _chk(probability, "probability")
    n = probability.shape[0]
    if out is None:
        out = torch.empty(n, dtype=probability.dtype, device=probability.device)
    _chk(out, "out", probability.dtype)
    f = _lib.fn("pisab_fill_probs", probability.dtype)
    _lib.check(f(_ptr(probability), int(initial_flav), int(flav), n, _ptr(out), _stream()))
    return out


def apply_osc_weights(nu_flux, prob_e, prob_mu, weights):
    """In place ``weights *= nu_flux[:,0]*prob_e + nu_flux[:,1]*prob_mu`` (prob3.py:621-622)."""
    _chk(weights, "weights")
    dt = weights.dtype
    n = weights.numel()
    _chk(nu_flux, "nu_flux", dt)
    _chk(prob_e, "prob_e", dt)
    _chk(prob_mu, "prob_mu", dt)
    if nu_flux.shape != (n, 2):
        raise ValueError("nu_flux must be [N, 2]")
    f = _lib.fn("pisab_apply_osc_weights", dt)
    _lib.check(f(_ptr(nu_flux), _ptr(prob_e), _ptr(prob_mu), n, _ptr(weights), _stream()))
    return weights


def scale_weights(weights, factor, scale):
    """In place ``weights *= factor * scale`` (aeff.aeff, pisa/stages/aeff/aeff.py:68-88); ``factor`` may be None."""
    _chk(weights, "weights")
    dt = weights.dtype
    _chk(factor, "factor", dt, allow_none=True)
    n = weights.numel()
    if factor is not None and factor.numel() != n:
        raise ValueError("factor and weights must have the same length")
    _lib.check(_lib.fn("pisab_scale_weights", dt)(_ptr(factor), float(scale), n, _ptr(weights), _stream()))
    return weights


# -------------------------------------------------------------------------- histogram -----

def joint_index(index_a, index_b, size_b, out=None):
    """Flat index on the joint binning a + b from the two sub-indices (-1 where either is outside)."""
    _chk(index_a, "index_a", torch.int32)
    _chk(index_b, "index_b", torch.int32)
    n = index_a.numel()
    if index_b.numel() != n:
        raise ValueError("the two indices must have the same length")
    if out is None:
        out = torch.empty(n, dtype=torch.int32, device=index_a.device)
    _chk(out, "out", torch.int32)
    _lib.check(_lib.load().pisab_joint_index(_ptr(index_a), _ptr(index_b), int(size_b), n, _ptr(out), _stream()))
    return out


def hist_transform(weights, unc, transform, want_errors=True):
    """``utils.hist`` with a binned calc_mode (hist.py:131-160): returns float64 (hist, sumw2, bin_unc2) with
    hist = (unc w) @ T etc.; ``unc`` may be None; the last two are None unless ``want_errors``."""
    _chk(weights, "weights")
    dt = weights.dtype
    _chk(unc, "unc", dt, allow_none=True)
    _chk(transform, "transform", dt)
    n_calc = weights.numel()
    if transform.dim() != 2 or transform.shape[0] != n_calc or (unc is not None and unc.numel() != n_calc):
        raise ValueError("transform must be [n_calc, n_out] with one weight per calc bin")
    n_out = transform.shape[1]
    mk = lambda: torch.empty(n_out, dtype=torch.float64, device=weights.device)  # noqa: E731
    h = mk()
    s2, b2 = (mk(), mk()) if want_errors else (None, None)
    _lib.check(_lib.fn("pisab_hist_transform", dt)(_ptr(weights), _ptr(unc), _ptr(transform), n_calc, n_out, _ptr(h),
                                                  _ptr(s2), _ptr(b2), _stream()))
    return h, s2, b2


def make_binning(dims, device):
    """dims: list of dicts {kind: 'lin'|'log'|'edges', n_bins, lo, hi, edges}.  For 'log', lo/hi
    are the RAW domain; its log (hist.py:118-120) is taken on the device by the library, with the function
    and precision used for the samples, so that samples equal to an edge stay on it.
    Returns (Binning struct, keep-alive list of edge tensors)."""
    if not 1 <= len(dims) <= _lib.MAX_DIMS:
        raise ValueError("1..%d dimensions supported" % _lib.MAX_DIMS)
    b = Binning()
    b.n_dims = len(dims)
    keep = []
    for i, d in enumerate(dims):
        kind = d["kind"]
        b.n_bins[i] = int(d["n_bins"])
        if kind == "edges":
            edges = torch.as_tensor(np.asarray(d["edges"], dtype=np.float64), device=device)
            if edges.numel() != b.n_bins[i] + 1:
                raise ValueError("edges must have n_bins + 1 entries")
            keep.append(edges)
            b.kind[i] = _lib.DIM_EDGES
            b.d_edges[i] = edges.data_ptr()
            b.lo[i], b.hi[i] = float(edges[0]), float(edges[-1])
        elif kind == "log":
            if not 0.0 < float(d["lo"]) < float(d["hi"]):
                raise ValueError("a logarithmic dimension needs 0 < lo < hi")
            b.kind[i], b.lo[i], b.hi[i] = _lib.DIM_LOG, float(d["lo"]), float(d["hi"])
        elif kind == "lin":
            b.kind[i], b.lo[i], b.hi[i] = _lib.DIM_LIN, float(d["lo"]), float(d["hi"])
        else:
            raise ValueError("unknown dimension kind %r" % kind)
    # the struct only holds raw device pointers: tie the edge tensors' lifetime to it, so that a caller that drops the
    # second return value cannot leave `d_edges` dangling
    b._keep = keep
    return b, keep


def hist_index(binning, coords, out=None):
    """Flat row-major bin index per event (int32, -1 outside); translation.py:417-455 rule."""
    dt = coords[0].dtype
    n = coords[0].numel()
    if len(coords) != binning.n_dims:
        raise ValueError("one sample array per binning dimension")
    for c in coords:
        _chk(c, "coords", dt)
        if c.numel() != n:
            raise ValueError("sample arrays must have equal length")
    if out is None:
        out = torch.empty(n, dtype=torch.int32, device=coords[0].device)
    _chk(out, "out", torch.int32)
    ptrs = (ctypes.c_void_p * len(coords))(*[c.data_ptr() for c in coords])
    f = _lib.fn("pisab_hist_index", dt)
    _lib.check(f(ctypes.byref(binning), ptrs, n, _ptr(out), _stream()))
    return out


_workspaces = {}


def _workspace(device, n, n_bins, n_containers=1):
    need = int(_lib.load().pisab_reweight_batch_workspace_bytes(n_containers, n_bins))
    key = (device.index if device.index is not None else torch.cuda.current_device())
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(max(need, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


class HistPlan:
    """Setup-time plan of a histogram over static bin indices; ``hist_accumulate(..., plan=plan)`` then needs only the
    current weights.  Up to DET_MAX_BINS (1024) bins (``pisab_hist_plan_build``): per tile of 2048 events the
    permutation grouping the events by bin and the group offsets.  Above: the SORTED plan (``perm`` = stable order of
    all events by bin, ``sorted_index`` = index[perm]; ``pisab_hist_accumulate_sorted``)."""

    def __init__(self, buf, n, n_bins, index, perm=None, sorted_index=None):
        self.buf, self.n, self.n_bins = buf, int(n), int(n_bins)
        self.perm, self.sorted_index = perm, sorted_index
        self._index, self._counts = index, None

    @property
    def counts(self):
        """Event counts per bin (float64): static like the plan, computed once."""
        if self._counts is None:
            self._counts = hist_accumulate(self._index, None, self.n_bins, want_w2=False)[0]
        return self._counts


def hist_plan(index, n_bins):
    _chk(index, "index", torch.int32)
    n = index.numel()
    if int(n_bins) > _lib.DET_MAX_BINS:
        # out-of-range events share the key n_bins: they sort to the end and the kernel skips them
        key = torch.where((index >= 0) & (index < n_bins), index, torch.full_like(index, int(n_bins)))
        perm, sorted_key = sort_order(key.contiguous(), want_sorted=True)
        return HistPlan(None, n, n_bins, index, perm=perm, sorted_index=sorted_key)
    nbytes = int(_lib.load().pisab_hist_plan_bytes(n, int(n_bins)))
    if nbytes == 0:
        return None
    buf = torch.empty(nbytes, dtype=torch.uint8, device=index.device)
    _lib.check(_lib.load().pisab_hist_plan_build(_ptr(index), n, int(n_bins), _ptr(buf), nbytes, _stream()))
    return HistPlan(buf, n, n_bins, index)


def hist_accumulate(index, weights, n_bins, want_w2=True, plan=None):
    """(hist, hist_w2) as float64 tensors; ``weights`` may be None (counts, hist.py:179-185).  With a ``plan``
    (``hist_plan`` of the same index) the bin-sorted-tile kernel is used."""
    _chk(index, "index", torch.int32)
    n = index.numel()
    if plan is not None and (plan.n != n or plan.n_bins != int(n_bins)):
        raise ValueError("the plan was built for %d events / %d bins" % (plan.n, plan.n_bins))
    if plan is not None and weights is None:
        c = plan.counts                      # unweighted histogram of static indices: nothing to recompute
        return c.clone(), (c.clone() if want_w2 else None)
    if plan is not None and weights is not None and plan.perm is not None:
        _chk(weights, "weights")
        if weights.numel() != n:
            raise ValueError("weights and index must have the same length")
        hist = torch.empty(n_bins, dtype=torch.float64, device=index.device)
        w2 = torch.empty(n_bins, dtype=torch.float64, device=index.device) if want_w2 else None
        ws = _workspace(index.device, n, n_bins)
        f = _lib.fn("pisab_hist_accumulate_sorted", weights.dtype)
        _lib.check(f(_ptr(plan.perm), _ptr(plan.sorted_index), _ptr(weights), n, int(n_bins), _ptr(hist), _ptr(w2),
                     _ptr(ws), ws.numel(), _stream()))
        return hist, w2
    if plan is not None and weights is not None and weights.data_ptr() % 16 == 0:
        _chk(weights, "weights")
        if weights.numel() != n:
            raise ValueError("weights and index must have the same length")
        hist = torch.empty(n_bins, dtype=torch.float64, device=index.device)
        w2 = torch.empty(n_bins, dtype=torch.float64, device=index.device) if want_w2 else None
        ws = _workspace(index.device, n, n_bins)
        f = _lib.fn("pisab_hist_accumulate_planned", weights.dtype)
        _lib.check(f(_ptr(plan.buf), _ptr(weights), n, int(n_bins), _ptr(hist), _ptr(w2), _ptr(ws), ws.numel(), _stream()))
        return hist, w2
    dt = torch.float64 if weights is None else weights.dtype
    if weights is not None:
        _chk(weights, "weights")
        if weights.numel() != n:
            raise ValueError("weights and index must have the same length")
    hist = torch.empty(n_bins, dtype=torch.float64, device=index.device)
    w2 = torch.empty(n_bins, dtype=torch.float64, device=index.device) if want_w2 else None
    ws = _workspace(index.device, n, n_bins)
    f = _lib.fn("pisab_hist_accumulate", dt)
    _lib.check(f(_ptr(index), _ptr(weights), n, int(n_bins), _ptr(hist), _ptr(w2), _ptr(ws), ws.numel(), _stream()))
    return hist, w2


def lookup(index, flat_hist, out=None):
    """``lookup`` on a regularised binning (translation.py:417-501): out = flat_hist[index], 0 outside."""
    _chk(index, "index", torch.int32)
    _chk(flat_hist, "flat_hist")
    n = index.numel()
    width = 1 if flat_hist.dim() == 1 else flat_hist.shape[1]
    shape = (n,) if flat_hist.dim() == 1 else (n, width)
    if out is None:
        out = torch.empty(shape, dtype=flat_hist.dtype, device=index.device)
    _chk(out, "out", flat_hist.dtype)
    f = _lib.fn("pisab_lookup", flat_hist.dtype)
    _lib.check(f(_ptr(index), _ptr(flat_hist), n, int(width), _ptr(out), _stream()))
    return out


def reweight_hist(consts, earth, nubar, flav, energy, coszen, nu_flux, weights_in, index, n_bins,
                  weights_out=None, prob_e=None, prob_mu=None, hist=None, hist_w2=None, order=None):
    """Fused template evaluation: prob3 + ``weights *= flux.prob`` + weighted histogram (w, w^2)."""
    _chk(energy, "energy")
    dt = energy.dtype
    n = energy.numel()
    for t, nm in ((coszen, "coszen"), (nu_flux, "nu_flux"), (weights_in, "weights_in")):
        _chk(t, nm, dt)
    _chk(index, "index", torch.int32)
    if nu_flux.shape != (n, 2) or coszen.numel() != n or weights_in.numel() != n or index.numel() != n:
        raise ValueError("inconsistent event array shapes")
    for t, nm in ((weights_out, "weights_out"), (prob_e, "prob_e"), (prob_mu, "prob_mu")):
        _chk(t, nm, dt, allow_none=True)
    nb, d_nb = _species(nubar, n, "nubar")
    fl, d_fl = _species(flav, n, "flav")
    if hist is None:
        hist = torch.empty(n_bins, dtype=torch.float64, device=energy.device)
    if hist_w2 is None:
        hist_w2 = torch.empty(n_bins, dtype=torch.float64, device=energy.device)
    _check_order(order, n)
    ws = _workspace(energy.device, n, n_bins)
    f = _lib.fn("pisab_reweight_hist", dt)
    _lib.check(f(ctypes.byref(consts), ctypes.byref(earth), nb, _ptr(d_nb), fl, _ptr(d_fl), _ptr(energy),
                 _ptr(coszen), _ptr(nu_flux), _ptr(weights_in), _ptr(index), _ptr(order), n, int(n_bins), _ptr(hist),
                 _ptr(hist_w2), _ptr(weights_out), _ptr(prob_e), _ptr(prob_mu), _ptr(ws), ws.numel(), _stream()))
    return hist, hist_w2


class TemplateBatch:
    """Descriptor array for ``reweight_hist_batch``: up to MAX_BATCH flavour containers evaluated in ONE
    kernel launch.  Built once (the event arrays of a fit do not move); ``scale`` (the aeff.aeff
    per-container factor) can be updated between templates with ``set_scale``."""

    def __init__(self, containers, n_bins):
        """containers: list of dicts with nubar, flav, energy, coszen, nu_flux, weights, index and optional
        order, weights_out, astro_weights (additive per-event term, hist.py:141-145), scale, flags; with ``flux_terms``, ``nu_flux_nominal`` and ``nubar_flux_nominal`` given
        (and ``flags | _lib.CONTAINER_FLUX_SYS``) the kernel evaluates flux.barr_simple itself and never reads
        ``nu_flux`` (which may then be None)."""
        if not 1 <= len(containers) <= MAX_BATCH:
            raise ValueError("a batch holds 1..%d containers" % MAX_BATCH)
        self.n_bins = int(n_bins)
        self.n = len(containers)
        self.desc = (ContainerDesc * self.n)()
        self._keep = []
        dt = None
        for d, c in zip(self.desc, containers):
            e = _chk(c["energy"], "energy")
            dt = dt or e.dtype
            n = e.numel()
            flags = int(c.get("flags", 0))
            fold = bool(flags & _lib.CONTAINER_FLUX_SYS)
            for nm in ("coszen", "weights"):
                _chk(c[nm], nm, dt)
            _chk(c.get("nu_flux"), "nu_flux", dt, allow_none=fold)
            _chk(c["index"], "index", torch.int32)
            if (c.get("nu_flux") is not None and c["nu_flux"].shape != (n, 2)) or c["coszen"].numel() != n \
                    or c["weights"].numel() != n or c["index"].numel() != n:
                raise ValueError("inconsistent event array shapes")
            terms, nom, nom_bar = c.get("flux_terms"), c.get("nu_flux_nominal"), c.get("nubar_flux_nominal")
            if fold:
                _chk(terms, "flux_terms", torch.float64)
                _chk(nom, "nu_flux_nominal", dt)
                _chk(nom_bar, "nubar_flux_nominal", dt)
                if terms.shape != (n, 4) or nom.shape != (n, 2) or nom_bar.shape != (n, 2):
                    raise ValueError("flux_terms must be [n, 4], the nominal fluxes [n, 2]")
            order, wout, astro = c.get("order"), c.get("weights_out"), c.get("astro_weights")
            _check_order(order, n)
            _chk(wout, "weights_out", dt, allow_none=True)
            _chk(astro, "astro_weights", dt, allow_none=True)
            if astro is not None and astro.numel() != n:
                raise ValueError("inconsistent event array shapes")
            d.d_energy, d.d_coszen = e.data_ptr(), c["coszen"].data_ptr()
            d.d_nu_flux = 0 if c.get("nu_flux") is None else c["nu_flux"].data_ptr()
            d.d_weights = c["weights"].data_ptr()
            d.d_index = c["index"].data_ptr()
            if fold:
                d.d_flux_terms, d.d_nu_flux_nominal, d.d_nubar_flux_nominal = terms.data_ptr(), nom.data_ptr(), nom_bar.data_ptr()
            d.d_order = 0 if order is None else order.data_ptr()
            d.d_weights_out = 0 if wout is None else wout.data_ptr()
            d.d_astro_weights = 0 if astro is None else astro.data_ptr()
            d.n, d.scale, d.nubar, d.flav = n, float(c.get("scale", 1.0)), int(c["nubar"]), int(c["flav"])
            d.flags = flags
            self._keep.append((e, c["coszen"], c.get("nu_flux"), c["weights"], c["index"], order, wout, terms, nom, nom_bar, astro))
        self.dtype = dt
        self.device = containers[0]["energy"].device

    def set_scale(self, i, scale):
        self.desc[i].scale = float(scale)


def flux_sys(nue_numu_ratio=1.0, nu_nubar_ratio=1.0, delta_index=0.0, Barr_uphor_ratio=0.0, Barr_nu_nubar_ratio=0.0):
    """pisab_flux_sys_t from the parameter names of flux.barr_simple (barr_simple.py:41-52)."""
    return _lib.FluxSys(float(nue_numu_ratio), float(nu_nubar_ratio), float(delta_index), float(Barr_uphor_ratio),
                        float(Barr_nu_nubar_ratio))


def _sys_ref(sys):
    return None if sys is None else ctypes.byref(sys)


def reweight_hist_batch(consts, earth, batch, out=None, flux_sys=None):
    """All containers of ``batch`` in one launch; returns ``[n_containers, 2, n_bins]`` (sum w, sum w^2).
    ``flux_sys`` (see ``flux_sys()``): the systematics for containers flagged CONTAINER_FLUX_SYS."""
    if out is None:
        out = torch.empty((batch.n, 2, batch.n_bins), dtype=torch.float64, device=batch.device)
    _chk(out, "out", torch.float64)
    if out.numel() != batch.n * 2 * batch.n_bins:
        raise ValueError("out must hold [n_containers, 2, n_bins] doubles")
    ws = _workspace(batch.device, 0, batch.n_bins, batch.n)
    f = _lib.fn("pisab_reweight_hist_batch", batch.dtype)
    _lib.check(f(ctypes.byref(consts), ctypes.byref(earth), batch.desc, batch.n, batch.n_bins, _sys_ref(flux_sys),
                 _ptr(out), _ptr(ws), ws.numel(), _stream()))
    return out


def reweight_hist_chi2(consts, earth, batch, observed, out=None, chi2=None, total=None, bin_scales=None,
                       flux_sys=None):
    """One hypothesis of a fit in one call: all containers of ``batch`` in one launch, then ONE kernel that reduces
    the partial histograms, applies optional per-bin scales (``bin_scales`` [n_containers, n_bins], the
    discr_sys.hypersurfaces factors), sums the containers and evaluates ``mod_chi2`` against ``observed`` [n_bins]
    into ``chi2`` (1-element float64 tensor or a slot of a scan's result array).  Returns (hist, chi2)."""
    if out is None:
        out = torch.empty((batch.n, 2, batch.n_bins), dtype=torch.float64, device=batch.device)
    _chk(out, "out", torch.float64)
    if out.numel() != batch.n * 2 * batch.n_bins:
        raise ValueError("out must hold [n_containers, 2, n_bins] doubles")
    _chk(observed, "observed", torch.float64, allow_none=True)
    if observed is not None and observed.numel() != batch.n_bins:
        raise ValueError("observed must be [n_bins]")
    if chi2 is None and observed is not None:
        chi2 = torch.empty(1, dtype=torch.float64, device=batch.device)
    _chk(chi2, "chi2", torch.float64, allow_none=True)
    _chk(total, "total", torch.float64, allow_none=True)
    _chk(bin_scales, "bin_scales", torch.float64, allow_none=True)
    if bin_scales is not None and bin_scales.numel() != batch.n * batch.n_bins:
        raise ValueError("bin_scales must be [n_containers, n_bins]")
    ws = _workspace(batch.device, 0, batch.n_bins, batch.n)
    f = _lib.fn("pisab_reweight_hist_chi2", batch.dtype)
    _lib.check(f(ctypes.byref(consts), ctypes.byref(earth), batch.desc, batch.n, batch.n_bins, _sys_ref(flux_sys),
                 _ptr(bin_scales), _ptr(observed), _ptr(out), _ptr(total), _ptr(chi2), _ptr(ws), ws.numel(), _stream()))
    return out, chi2


_scan_ws = {}


def reweight_hist_scan(consts_list, earth, batch, out=None):
    """P templates (``consts_list``: sequence of OscConsts) over the containers of ``batch`` in ONE launch;
    returns ``[P, n_containers, 2, n_bins]``."""
    n_t = len(consts_list)
    if n_t < 1:
        raise ValueError("at least one template")
    arr = consts_list if isinstance(consts_list, ctypes.Array) else (OscConsts * n_t)(*consts_list)
    if out is None:
        out = torch.empty((n_t, batch.n, 2, batch.n_bins), dtype=torch.float64, device=batch.device)
    _chk(out, "out", torch.float64)
    if out.numel() != n_t * batch.n * 2 * batch.n_bins:
        raise ValueError("out must hold [n_templates, n_containers, 2, n_bins] doubles")
    n_max = max(int(d.n) for d in batch.desc)
    need = int(_lib.load().pisab_reweight_scan_workspace_bytes(n_t, batch.n, batch.n_bins, n_max))
    key = batch.device.index if batch.device.index is not None else torch.cuda.current_device()
    ws = _scan_ws.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=batch.device)
        _scan_ws[key] = ws
    f = _lib.fn("pisab_reweight_hist_scan", batch.dtype)
    _lib.check(f(arr, n_t, ctypes.byref(earth), batch.desc, batch.n, batch.n_bins, _ptr(out), _ptr(ws), ws.numel(),
                 _stream()))
    return out


def template_chi2_batch(hist, observed, out=None):
    """One ``mod_chi2`` per template: ``hist`` is ``[P, n_containers, 2, n_bins]``; returns float64 ``[P]``."""
    _chk(hist, "hist", torch.float64)
    _chk(observed, "observed", torch.float64)
    if hist.dim() != 4 or hist.shape[2] != 2 or observed.numel() != hist.shape[3]:
        raise ValueError("hist must be [n_templates, n_containers, 2, n_bins] and observed [n_bins]")
    if out is None:
        out = torch.empty(hist.shape[0], dtype=torch.float64, device=hist.device)
    _chk(out, "out", torch.float64)
    _lib.check(_lib.load().pisab_template_chi2_batch(_ptr(hist), hist.shape[0], hist.shape[1], hist.shape[3],
                                                     _ptr(observed), _ptr(out), _stream()))
    return out


def mod_chi2(expected, expected_w2, observed):
    """``mod_chi2`` (pisa/utils/stats.py:651-695) on device; returns a 1-element float64 tensor."""
    _chk(expected, "expected", torch.float64)
    _chk(observed, "observed", torch.float64)
    _chk(expected_w2, "expected_w2", torch.float64, allow_none=True)
    out = torch.empty(1, dtype=torch.float64, device=expected.device)
    _lib.check(_lib.load().pisab_mod_chi2(_ptr(expected), _ptr(expected_w2), _ptr(observed), expected.numel(),
                                          _ptr(out), _stream()))
    return out


def template_chi2(hist, observed, out=None, total=None):
    """``mod_chi2`` of one template against ``observed``: ``hist`` is ``[n_containers, 2, n_bins]`` as returned
    by ``reweight_hist_batch``; containers are summed (MapSet sum, sumw2 errors).  ``out``: 1-element
    float64 tensor (or a view into a scan's result array); ``total``: optional ``[2, n_bins]`` output."""
    _chk(hist, "hist", torch.float64)
    _chk(observed, "observed", torch.float64)
    if hist.dim() != 3 or hist.shape[1] != 2 or observed.numel() != hist.shape[2]:
        raise ValueError("hist must be [n_containers, 2, n_bins] and observed [n_bins]")
    if out is None:
        out = torch.empty(1, dtype=torch.float64, device=hist.device)
    _chk(out, "out", torch.float64)
    _chk(total, "total", torch.float64, allow_none=True)
    _lib.check(_lib.load().pisab_template_chi2(_ptr(hist), hist.shape[0], hist.shape[2], _ptr(observed), _ptr(total),
                                               _ptr(out), _stream()))
    return out


def hist_reduce_chi2(partials, n_blocks, bin_scales=None, observed=None, total=None, chi2=None):
    """The fit-loop epilogue on its own (``pisab_hist_scale_sum_chi2``): ``partials`` [n_containers, n_blocks, 2, n_bins]
    (or [n_containers, 2, n_bins] with ``n_blocks`` = 1) summed in block order, per-bin ``bin_scales`` [n_containers,
    n_bins] applied like discr_sys.hypersurfaces, containers summed into ``total`` [2, n_bins], ``mod_chi2`` against
    ``observed`` into ``chi2``.  Returns the scaled per-container histograms [n_containers, 2, n_bins]."""
    _chk(partials, "partials", torch.float64)
    n_c, n_bins = partials.shape[0], partials.shape[-1]
    if partials.numel() != n_c * int(n_blocks) * 2 * n_bins:
        raise ValueError("partials must be [n_containers, n_blocks, 2, n_bins]")
    for t, name, numel in ((bin_scales, "bin_scales", n_c * n_bins), (observed, "observed", n_bins),
                           (total, "total", 2 * n_bins), (chi2, "chi2", 1)):
        _chk(t, name, torch.float64, allow_none=True)
        if t is not None and t.numel() != numel:
            raise ValueError("%s has the wrong size" % name)
    if observed is not None and chi2 is None:
        chi2 = torch.empty(1, dtype=torch.float64, device=partials.device)
    out = torch.empty((n_c, 2, n_bins), dtype=torch.float64, device=partials.device)
    _lib.check(_lib.load().pisab_hist_scale_sum_chi2(_ptr(partials), int(n_blocks), n_c, n_bins, _ptr(bin_scales),
                                                     _ptr(observed), _ptr(out), _ptr(total), _ptr(chi2), _stream()))
    return out


def fp64_peak_probe(iters=20000):
    """Measured DFMA throughput of the device (FLOP/s): the FP64 roofline denominator."""
    flops, ms = ctypes.c_double(), ctypes.c_double()
    _lib.check(_lib.load().pisab_fp64_peak_probe(int(iters), ctypes.byref(flops), ctypes.byref(ms)))
    return flops.value, ms.value


def launch_count(reset=False):
    return int(_lib.load().pisab_launch_count(1 if reset else 0))
