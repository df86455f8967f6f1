"""Template-evaluation engine: the fit-loop fast path over a set of event containers.

One call = one template: for every container (nue_cc ... nutaubar_nc, prob3.py:401-404) run the
fused kernel (prob3 through the Earth -> ``weights *= flux.prob`` -> weighted histogram with
sumw2) and return one device buffer ``[n_containers, 2, n_bins]`` (sum w, sum w^2).  With
``torch.distributed`` initialised, events are sharded over ranks and the buffer -- the only
thing that crosses NVLink -- is all-reduced once per template.

This is what ``Pipeline.run()`` of ``osc.prob3`` + ``utils.hist`` in events mode computes
(prob3.py:452-622, hist.py:129-218) without the intermediate HBM round trips.
"""
import numpy as np
import torch

from . import _lib, ops
from .distributed import combine_histograms, event_sharding, shard_slice  # noqa: F401  (shard_slice re-exported)

_NP2T = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32}

EVENT_KEYS = ("true_energy", "true_coszen", "nu_flux", "weights", "index")


class _Block:
    __slots__ = ("name", "nubar", "flav", "n", "n_real", "dev", "host", "stage", "order", "pack_index", "flags")

    def __init__(self, name, nubar, flav, n):
        self.name, self.nubar, self.flav, self.n = name, int(nubar), int(flav), int(n)
        self.n_real = self.n                 # without the zero-weight padding of the pair-aligned layout
        self.dev, self.host, self.stage, self.order = {}, {}, None, None
        self.pack_index = False
        self.flags = 0


class ReweightEngine:
    """Device-resident event containers + fused template evaluation."""

    def __init__(self, earth, n_bins, dtype=np.float64, device=None, sort_events=True, deterministic=True):
        if not torch.cuda.is_available():
            raise RuntimeError("pisa_b200.engine needs a CUDA device (there is no CPU path)")
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.earth = earth
        self.n_bins = int(n_bins)
        self.tdtype = _NP2T[np.dtype(dtype)]
        self.blocks = []
        self.sort_events = bool(sort_events)
        self.deterministic = bool(deterministic)
        self._out = None
        self._copy_stream = None
        self._host_out = None
        self._batches = None
        self._host_batches = {}

    # ------------------------------------------------------------------ containers -------
    def add_container(self, name, nubar, flav, true_energy, true_coszen, nu_flux, weights, index,
                      nu_flux_nominal=None, nubar_flux_nominal=None, astro_weights=None):
        """Register one container.  Tensors may live on the device (resident mode) or be pinned
        host tensors / numpy arrays (host mode, see evaluate_host).  With the nominal fluxes given (device
        tensors [n, 2]) the Barr flux systematics can float in the fit loop (``set_flux_params``: flux.barr_simple
        evaluated inside the template kernel, or ``nu_flux`` rewritten from them in one HBM-bound pass).
        ``astro_weights`` (device tensor [n]): the additive per-event term of utils.hist (hist.py:141-145),
        ``w = weights * (flux . prob) * scale + astro_weights``."""
        arrays = dict(true_energy=true_energy, true_coszen=true_coszen, nu_flux=nu_flux, weights=weights,
                      index=index)
        if (nu_flux_nominal is None) != (nubar_flux_nominal is None):
            raise ValueError("nu_flux_nominal and nubar_flux_nominal go together")
        if nu_flux_nominal is not None:
            arrays.update(nu_flux_nominal=nu_flux_nominal, nubar_flux_nominal=nubar_flux_nominal)
        if astro_weights is not None:
            if not torch.as_tensor(astro_weights).is_cuda:
                raise NotImplementedError("astro_weights: resident mode only")
            arrays.update(astro_weights=astro_weights)
        n = int(arrays["true_energy"].shape[0])
        blk = _Block(name, nubar, flav, n)
        for k, a in arrays.items():
            want = torch.int32 if k == "index" else self.tdtype
            t = torch.as_tensor(a)
            if t.dtype != want:
                raise TypeError("%s.%s must be %s" % (name, k, want))
            if t.is_cuda:
                blk.dev[k] = t.contiguous()
            else:
                blk.host[k] = t.contiguous().pin_memory() if not t.is_pinned() else t
        blk.pack_index = "index" in blk.host and self.n_bins < 255
        if self.sort_events:
            # setup-time, depends on true_coszen only (like calcLayers in prob3.setup_function): the events
            # are physically re-ordered so that a warp's 32 events cross the same number of Earth shells
            # (no divergence) AND its loads are contiguous; a histogram does not depend on the event order
            cz = blk.dev["true_coszen"] if "true_coszen" in blk.dev else blk.host["true_coszen"].to(self.device)
            dummy = None
            # more bins than fit in shared memory: the events of a class are ordered by bin as well, so that a warp's
            # 32 events share a bin and cost one exact atomic add instead of 32 (warp_fixed_add)
            by_bin = None
            if self.n_bins > _lib.DET_MAX_BINS:
                by_bin = blk.dev["index"] if "index" in blk.dev else blk.host["index"].to(self.device)
            if self.tdtype == torch.float32 and n > 0:
                # FP32 mode: the template kernel handles TWO events per thread in the lanes of the packed FP32
                # instructions; pairs must cross the same shells, so every class is padded to an even size with a
                # repeat of its last event carrying weight 0 and bin -1 (at most one per class)
                order, dummy = ops.pair_aligned_order(self.earth, cz, by_bin)
                blk.flags |= _lib.CONTAINER_PAIR_ALIGNED
                blk.n = int(order.numel())
            else:
                order = ops.layer_order(self.earth, cz, by_bin).long()

            def rearranged(t, key, idx, pad):
                out = t[idx].contiguous()
                if pad is not None and key in ("weights", "index", "astro_weights"):
                    out[pad] = -1 if key == "index" else 0
                return out
            for k in list(blk.dev):
                blk.dev[k] = rearranged(blk.dev[k], k, order, dummy)
            if blk.host:
                order_h, dummy_h = order.cpu(), None if dummy is None else dummy.cpu()
                for k in list(blk.host):
                    blk.host[k] = rearranged(blk.host[k], k, order_h, dummy_h).pin_memory()
        if blk.pack_index:
            # host mode: the static bin index (-1 .. n_bins-1) travels as index + 1 in ONE byte per event and is
            # widened on the device after the copy (41 instead of 44 bytes per event over PCIe)
            idx = blk.host.pop("index")
            idx = torch.where((idx >= 0) & (idx < self.n_bins), idx, torch.full_like(idx, -1))   # anything else is "outside"
            blk.host["index_u8"] = (idx + 1).to(torch.uint8).contiguous().pin_memory()
        if "nu_flux_nominal" in blk.dev:
            # parameter-independent terms of flux.barr_simple, once (after the re-ordering)
            blk.dev["flux_barr_terms"] = ops.flux_barr_terms(blk.dev["true_energy"], blk.dev["true_coszen"])
        self.blocks.append(blk)
        self._out = None
        self._batches = None
        self._batches_fold = None
        self._flux_batches = None
        return blk

    @property
    def n_events(self):
        return sum(b.n_real for b in self.blocks)

    def _result_buffer(self):
        if self._out is None or self._out.shape[0] != len(self.blocks):
            self._out = torch.empty((len(self.blocks), 2, self.n_bins), dtype=torch.float64, device=self.device)
        return self._out

    # ------------------------------------------------------------------- evaluation ------
    def set_flux_params(self, nue_numu_ratio=1.0, nu_nubar_ratio=1.0, delta_index=0.0, Barr_uphor_ratio=0.0,
                        Barr_nu_nubar_ratio=0.0, materialize=None):
        """flux.barr_simple for every container registered with nominal fluxes.  ``materialize=False``: nothing is
        launched; the five parameters ride along with the next ``evaluate`` / ``evaluate_chi2``, whose template kernel
        evaluates the systematics per event in registers (PISAB_CONTAINER_FLUX_SYS) -- ``nu_flux`` is neither written
        nor read.  ``materialize=True`` (and every path that needs the array: ``evaluate_many``, ``astro_weights``,
        more than ``DET_MAX_BINS`` bins) rewrites ``nu_flux`` in place instead (``pisab_flux_barr_apply_batch``, one
        launch for all containers, 80 B/event at ~90 % of the HBM roofline).  Both forms give the same bits.
        Default (None): inside the kernel for float64 (FP64-bound kernel: 11.93 instead of 12.30 ms per 1e8 events)
        and for samples below 1e6 events (one launch less per template); as a separate pass for large float32 samples,
        whose issue-bound kernel pays more for the FP64 flux arithmetic than the HBM-bound pass costs (8.33 vs
        8.26 ms; scratch/bench_flux_fold.py)."""
        for blk in self.blocks:
            if "flux_barr_terms" not in blk.dev:
                raise ValueError("container %s was registered without nominal fluxes" % blk.name)
        self._flux_sys = ops.flux_sys(nue_numu_ratio, nu_nubar_ratio, delta_index, Barr_uphor_ratio, Barr_nu_nubar_ratio)
        self._flux_stale = True
        if materialize is None:
            materialize = self.tdtype == torch.float32 and self.n_events >= 1_000_000
        if materialize or self.n_bins > _lib.DET_MAX_BINS:
            self._materialize_flux()

    def _materialize_flux(self):
        """Write ``nu_flux`` for the current flux systematics if it is out of date."""
        if not getattr(self, "_flux_stale", False):
            return
        if getattr(self, "_flux_batches", None) is None or self._flux_batches[0] != len(self.blocks):
            chunks = [self.blocks[lo:lo + ops.MAX_BATCH] for lo in range(0, len(self.blocks), ops.MAX_BATCH)]
            self._flux_batches = (len(self.blocks), [ops.FluxBatch([dict(
                terms=b.dev["flux_barr_terms"], nu_flux_nominal=b.dev["nu_flux_nominal"],
                nubar_flux_nominal=b.dev["nubar_flux_nominal"], nu_flux=b.dev["nu_flux"], nubar=b.nubar) for b in ch])
                for ch in chunks])
        y = self._flux_sys
        for fb in self._flux_batches[1]:      # ONE launch for all containers of a template
            ops.flux_barr_apply_batch(fb, y.nue_numu_ratio, y.nu_nubar_ratio, y.delta_index, y.barr_uphor_ratio,
                                      y.barr_nu_nubar_ratio)
        self._flux_stale = False

    def set_scales(self, scales):
        """Per-container factor folded into the weights (aeff.aeff: livetime * aeff_scale * norms)."""
        self.scales = [float(x) for x in scales]
        if len(self.scales) != len(self.blocks):
            raise ValueError("one scale per container")
        self._batches = None
        self._batches_fold = None
        self._host_batches = {}

    def _get_batches(self, fold_flux=False):
        """Descriptor arrays of the template launches; ``fold_flux``: the variant whose containers carry the cached
        flux terms and nominal fluxes instead of ``nu_flux`` (PISAB_CONTAINER_FLUX_SYS)."""
        key = "_batches_fold" if fold_flux else "_batches"
        if getattr(self, key, None) is None:
            scales = getattr(self, "scales", None) or [1.0] * len(self.blocks)
            batches = []
            for lo in range(0, len(self.blocks), ops.MAX_BATCH):
                chunk = self.blocks[lo:lo + ops.MAX_BATCH]
                desc = [dict(nubar=b.nubar, flav=b.flav, energy=b.dev["true_energy"], coszen=b.dev["true_coszen"],
                             nu_flux=b.dev["nu_flux"], weights=b.dev["weights"], index=b.dev["index"],
                             scale=scales[lo + j], flags=b.flags, astro_weights=b.dev.get("astro_weights"))
                        for j, b in enumerate(chunk)]
                if fold_flux:
                    for d, b in zip(desc, chunk):
                        d.update(flags=b.flags | _lib.CONTAINER_FLUX_SYS, flux_terms=b.dev["flux_barr_terms"],
                                 nu_flux_nominal=b.dev["nu_flux_nominal"], nubar_flux_nominal=b.dev["nubar_flux_nominal"])
                batches.append((lo, ops.TemplateBatch(desc, self.n_bins)))
            setattr(self, key, batches)
        return getattr(self, key)

    def _flux_fold(self):
        """(batches, flux_sys) of the next template launch: the folded form while ``nu_flux`` is out of date."""
        if getattr(self, "_flux_stale", False):
            if any("astro_weights" in b.dev for b in self.blocks):   # not a fit-loop (PLAIN) launch: write the array
                self._materialize_flux()
                return self._get_batches(), None
            return self._get_batches(fold_flux=True), self._flux_sys
        return self._get_batches(), None

    def evaluate(self, consts, allreduce=None, events=None):
        """Resident mode: all event arrays already in HBM.  ONE fused launch for all containers
        (+ one reduction of the per-block partial histograms).  Returns [n_containers, 2, n_bins].
        ``events``: optional list that receives one (start, stop) CUDA-event pair per fused
        launch (for the roofline timing in bench.py)."""
        allreduce = self._want_exchange(allreduce)
        out = self._result_buffer()
        # (binnings beyond _lib.DET_MAX_BINS take the same launch: the library switches to exact fixed-point
        # accumulators in global memory, bit-reproducible as well)
        batches, sys = self._flux_fold()
        for lo, batch in batches:
            if events is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            ops.reweight_hist_batch(consts, self.earth, batch, out=out[lo:lo + batch.n], flux_sys=sys)
            if events is not None:
                e1.record()
                events.append((e0, e1))
        if allreduce:
            self.allreduce(out)
        return out

    def evaluate_chi2(self, consts, observed, chi2_out=None, bin_scales=None, allreduce=None):
        """One hypothesis of a fit: template + ``mod_chi2`` against ``observed`` ([n_bins] float64 device tensor) with
        nothing synchronising.  On one rank this is ONE library call and two launches (``pisab_reweight_hist_chi2``:
        template kernel, then reduce + per-bin ``bin_scales`` + container sum + chi2 in one kernel); with events sharded
        over GPUs the exchange sits between the reduction and the chi2 (three launches).  Returns (hist, chi2)."""
        batches, sys = self._flux_fold()
        out = self._result_buffer()
        if chi2_out is None:
            chi2_out = torch.empty(1, dtype=torch.float64, device=self.device)
        if len(batches) == 1 and not self._want_exchange(allreduce) and self.n_bins <= _lib.DET_MAX_BINS:
            ops.reweight_hist_chi2(consts, self.earth, batches[0][1], observed, out=out, chi2=chi2_out,
                                   bin_scales=bin_scales, flux_sys=sys)
            return out, chi2_out
        out = self.evaluate(consts, allreduce=allreduce)
        if bin_scales is not None:
            out = out.clone()
            out[:, 0] = torch.clamp(out[:, 0] * bin_scales, min=0.0)
            out[:, 1] = out[:, 1] * bin_scales * bin_scales
        ops.template_chi2(out, observed, out=chi2_out)
        return out, chi2_out

    def evaluate_many(self, consts_list, allreduce=None):
        """P hypotheses in ONE launch (``pisab_reweight_hist_scan``): returns ``[P, n_containers, 2, n_bins]``.
        The single histogram exchange covers all P templates when sharded over GPUs."""
        self._materialize_flux()
        batches = self._get_batches()
        if len(batches) != 1:
            raise NotImplementedError("evaluate_many supports up to %d containers" % ops.MAX_BATCH)
        out = ops.reweight_hist_scan(consts_list, self.earth, batches[0][1])
        if self._want_exchange(allreduce):
            self.allreduce(out)
        return out

    def evaluate_host(self, consts, allreduce=None, changed=None):
        """Host mode: event arrays live in pinned host memory; every call copies them to the
        device (double-buffered on a copy stream so the copy of container i+1 overlaps the
        kernel of container i) and returns the histograms as a HOST numpy array.

        ``changed``: None = every array travels on every call (44 B/event in FP64, 41 with the packed bin index);
        or a tuple of keys (e.g. ``("weights", "nu_flux")``, the only arrays a fit changes between templates): the
        other arrays are uploaded ONCE on the first such call and stay resident, so a call moves 24 B/event."""
        if changed is not None:
            changed = tuple(changed)
            unknown = [k for k in changed if k not in EVENT_KEYS]
            if unknown:
                raise ValueError("unknown event arrays %s (known: %s)" % (unknown, ", ".join(EVENT_KEYS)))
            for blk in self.blocks:
                for k in EVENT_KEYS:
                    if k not in changed and k not in blk.dev:
                        src = blk.host["index_u8"] if (k == "index" and blk.pack_index) else blk.host[k]
                        t = src.to(self.device, non_blocking=True)
                        blk.dev[k] = (t.to(torch.int32) - 1) if (k == "index" and blk.pack_index) else t
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        out = self._result_buffer()
        main = torch.cuda.current_stream()
        nmax = max(b.n for b in self.blocks)
        stages = getattr(self, "_stages", None)
        if stages is None or stages[0]["true_energy"].shape[0] < nmax:
            stages = []
            for _ in range(2):
                stages.append(dict(
                    true_energy=torch.empty(nmax, dtype=self.tdtype, device=self.device),
                    true_coszen=torch.empty(nmax, dtype=self.tdtype, device=self.device),
                    nu_flux=torch.empty((nmax, 2), dtype=self.tdtype, device=self.device),
                    weights=torch.empty(nmax, dtype=self.tdtype, device=self.device),
                    index=torch.empty(nmax, dtype=torch.int32, device=self.device),
                    index_u8=torch.empty(nmax, dtype=torch.uint8, device=self.device)))
            self._stages = stages
            self._host_batches = {}
            self._stage_free = [torch.cuda.Event(), torch.cuda.Event()]
            self._stage_ready = [torch.cuda.Event(), torch.cuda.Event()]
            for e in self._stage_free:
                e.record(main)
        h2d = 0
        for i, blk in enumerate(self.blocks):
            s = i & 1
            st = stages[s]
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(self._stage_free[s])
                views = {}
                for k in EVENT_KEYS:
                    if changed is not None and k not in changed:
                        views[k] = blk.dev[k]          # resident static array
                        continue
                    packed = k == "index" and blk.pack_index
                    src = blk.host["index_u8" if packed else k]
                    dst = st[k][:blk.n]
                    if packed:
                        raw = st["index_u8"][:blk.n]
                        raw.copy_(src, non_blocking=True)
                        dst.copy_(raw)          # uint8 -> int32 on the device
                        dst.sub_(1)
                    else:
                        dst.copy_(src, non_blocking=True)
                    views[k] = dst
                    h2d += src.numel() * src.element_size()
                self._stage_ready[s].record(self._copy_stream)
            main.wait_event(self._stage_ready[s])
            batch = self._host_batches.get((i, s, changed))
            if batch is None:
                scales = getattr(self, "scales", None) or [1.0] * len(self.blocks)
                batch = ops.TemplateBatch([dict(nubar=blk.nubar, flav=blk.flav, energy=views["true_energy"],
                                                coszen=views["true_coszen"], nu_flux=views["nu_flux"],
                                                weights=views["weights"], index=views["index"],
                                                scale=scales[i], flags=blk.flags)], self.n_bins)
                self._host_batches[(i, s, changed)] = batch
            ops.reweight_hist_batch(consts, self.earth, batch, out=out[i:i + 1])
            self._stage_free[s].record(main)
        if self._want_exchange(allreduce):
            self.allreduce(out)
        if self._host_out is None or self._host_out.shape != out.shape:
            self._host_out = torch.empty(out.shape, dtype=out.dtype).pin_memory()
        self._host_out.copy_(out, non_blocking=True)
        main.synchronize()
        self.last_h2d_bytes = h2d
        self.last_d2h_bytes = out.numel() * out.element_size()
        return self._host_out.numpy()

    @staticmethod
    def _want_exchange(allreduce):
        """``allreduce=None`` (default) exchanges exactly when the events are sharded over the ranks
        (``pisa_b200.distributed.enable_event_sharding``): a process group initialised for another purpose must not
        silently sum histograms across ranks.  True / False force it."""
        return event_sharding() if allreduce is None else bool(allreduce)

    def allreduce(self, buf):
        """Sum the per-GPU histograms: the single exchange step per template (no-op on one rank)."""
        return combine_histograms(buf, deterministic=self.deterministic)
