"""``Container`` / ``VirtualContainer`` / ``ContainerSet`` with DEVICE-RESIDENT arrays.

Same contract as pisa/core/container.py (reference :199-1040): a container holds named variables
in several *representations* ("events", "log_events", any ``MultiDimBinning``), tracks per
variable which representations are valid (``validity[key][hash(rep)]``), and translates on demand
(``auto_translate`` :890-895):

    events  -> log_events   log of the sample                     (:845-850)
    events  -> binned       histogram, averaged or summed          (``array_to_binned`` :933-979)
    binned  -> events       lookup of the bin an event falls into  (``binned_to_array`` :981-1012)

What is different: arrays live in HBM as ``torch`` CUDA tensors (assigning a numpy array uploads
it once), and the two translations that touch every event run in the CUDA library
(``ops.hist_index`` + ``ops.hist_accumulate`` / ``ops.lookup``); the flat bin index of every
(container, binning) pair is computed once and cached, because the sample coordinates of a fit do
not change between templates.  ``get_map`` / ``get_mapset`` return host ``Map`` objects.
Binned -> binned goes through ``resample`` (average mode only, as in the reference).
"""
from collections import defaultdict
from collections.abc import Sequence

import numpy as np
import torch

from pisa_b200 import FTYPE
from pisa_b200.core.binning import MultiDimBinning, OneDimBinning
from pisa_b200.core.map import Map, MapSet

__all__ = ["Container", "VirtualContainer", "ContainerSet", "default_device", "TDTYPE"]

TDTYPE = torch.float64 if FTYPE == np.float64 else torch.float32


_DEVICES = {}


def default_device():
    # asked on every container write: cache one torch.device per CUDA ordinal
    if not _DEVICES and not torch.cuda.is_available():
        raise RuntimeError("pisa_b200 containers are device resident and need a CUDA device "
                           "(there is no CPU fallback)")
    ordinal = torch.cuda.current_device()
    dev = _DEVICES.get(ordinal)
    if dev is None:
        dev = _DEVICES[ordinal] = torch.device("cuda", ordinal)
    return dev


def regularized_dims(binning, policy="hist"):
    """Per-dimension description of a binning for ops.make_binning.

    policy "hist"   : what ``utils.hist`` does (hist.py:86-127): irregular -> explicit edges (searchsorted on the real
                      edges), log -> linear bins in log(x) (half-open, ``log_events`` sample), else linear (half-open);
    policy "generic": what the Container translations do (translation.histogram / lookup, translation.py:88-129,
                      228-344): a binning whose dimensions are ALL linear-regular goes through the fast_histogram /
                      lookup_regular rule (half-open); as soon as one dimension is logarithmic or irregular the WHOLE
                      binning goes through ``np.histogramdd`` / ``find_index`` on the real edges, i.e. the upper edge
                      of every dimension is inclusive;
    policy "edges"  : every dimension by its real edges (the second branch of "generic", forced: used for the two
                      halves of a joint binning)."""
    if policy not in ("hist", "generic", "edges"):
        raise ValueError("unknown index policy %r" % policy)
    if policy == "generic" and any(d.is_irregular or d.is_log for d in binning):
        policy = "edges"
    dims = []
    for d in binning:
        edges = np.asarray(d.bin_edges.magnitude, dtype=np.float64)
        if policy == "edges" or d.is_irregular:
            dims.append(dict(kind="edges", n_bins=d.num_bins, edges=edges))
        elif d.is_log:
            dims.append(dict(kind="log", n_bins=d.num_bins, lo=float(edges[0]), hi=float(edges[-1])))
        else:
            dims.append(dict(kind="lin", n_bins=d.num_bins, lo=float(edges[0]), hi=float(edges[-1])))
    return dims


class Container:
    valid_translation_modes = ("average", "sum")
    sum_mode_keys = ()
    array_representations = ("events", "log_events")

    def __init__(self, name, representation="events", device=None):
        self.name = name
        self.device = device
        self._representation = None
        self.linked = False
        self._aux_data = {}
        self.validity = defaultdict(dict)
        self.translation_modes = {}
        self.data = defaultdict(dict)
        self._representations = {}
        self.precedence = defaultdict(int)
        self._index_cache = {}
        self._index_cache_names = set()   # coordinate names any cached bin index depends on
        self.representation = representation

    def __repr__(self):
        return "Container containing keys %s" % self.all_keys

    # ----------------------------------------------------------------- representation ----
    @property
    def representation(self):
        return self._representation

    @representation.setter
    def representation(self, representation):
        key = hash(representation)
        if key not in self._representations:
            self._representations[key] = representation
            if isinstance(representation, MultiDimBinning):
                for name in representation.names:
                    self.validity[name][key] = True
            elif isinstance(representation, str):
                if representation not in self.array_representations:
                    raise ValueError("Unknown representation '%s'" % representation)
        self._representation = representation
        self.current_data = self.data[key]

    representations = property(lambda self: tuple(self._representations.values()))
    representation_keys = property(lambda self: tuple(self._representations.keys()))
    is_map = property(lambda self: isinstance(self._representation, MultiDimBinning))

    def set_aux_data(self, key, val):
        if key in self.all_keys:
            raise KeyError("Key %s already exsits" % key)
        self._aux_data[key] = val

    @property
    def shape(self):
        if self.is_map:
            return self.representation.shape
        if len(self.keys) == 0:
            return None
        return tuple(self[self.keys[0]].shape[0:1])

    @property
    def size(self):
        shape = self.shape
        return int(np.prod(shape)) if shape is not None else 0

    @property
    def num_dims(self):
        return self.representation.num_dims if self.is_map else 1

    @property
    def keys(self):
        keys = tuple(self.current_data.keys())
        if self.is_map:
            keys += tuple(self.representation.names)
        return keys

    keys_incl_aux_data = property(lambda self: list(self.keys) + list(self._aux_data.keys()))
    all_keys = property(lambda self: list(self.validity.keys()))
    all_keys_incl_aux_data = property(lambda self: self.all_keys + list(self._aux_data.keys()))

    # ------------------------------------------------------------------------ validity ---
    def mark_changed(self, key):
        """Only the copy in the current representation stays valid (container.py:638-649)."""
        for rep in self.validity[key]:
            self.validity[key][rep] = False
        if key in self.current_data.keys():
            self.mark_valid(key)
        # a changed coordinate invalidates every cached bin index that used it
        if key in self._index_cache_names:
            for ck in [ck for ck, (names, _) in self._index_cache.items() if key in names]:
                del self._index_cache[ck]

    def mark_valid(self, key):
        self.validity[key][hash(self.representation)] = True

    # -------------------------------------------------------------------------- access ---
    def _to_device(self, data):
        dev = self.device or default_device()
        if isinstance(data, torch.Tensor):
            if data.device == dev and (data.dtype == TDTYPE or not data.is_floating_point()) and data.is_contiguous():
                return data
            t = data.to(dev)
        else:
            t = torch.as_tensor(np.ascontiguousarray(data), device=dev)
        if t.is_floating_point() and t.dtype != TDTYPE:
            t = t.to(TDTYPE)
        return t.contiguous()

    def __setitem__(self, key, data):
        if self.is_map and key in self.representation.names:
            raise Exception("Cannot add variable %s, as it is a binning dimension" % key)
        self.__add_data(key, data)
        if key not in self.translation_modes:
            self.translation_modes[key] = "sum" if key in self.sum_mode_keys else "average"
        self.mark_changed(key)

    def __add_data(self, key, data):
        if isinstance(data, (np.ndarray, torch.Tensor)):
            if self.is_map:
                self.__add_data(key, (self.representation, data))
            else:
                shape = self.shape
                if shape is not None:
                    assert tuple(data.shape[:self.num_dims]) == shape, "Incompatible dimensions"
                self.current_data[key] = self._to_device(data)
        elif isinstance(data, Map):
            assert hash(self.representation) == hash(data.binning)
            self.current_data[key] = self._to_device(data.hist.ravel())
        elif isinstance(data, Sequence) and len(data) == 2:
            binning, array = data
            assert isinstance(binning, MultiDimBinning)
            assert hash(self.representation) == hash(binning)
            if array.shape[0] == binning.size:
                flat = array
            else:
                assert tuple(array.shape[:binning.num_dims]) == binning.shape
                flat = array.reshape((binning.size, -1) if array.ndim > binning.num_dims else (binning.size,))
            self.current_data[key] = self._to_device(flat)
        else:
            raise TypeError("unknown dataformat")

    def __getitem__(self, key):
        if self.is_map:
            binning = self.representation
            if key in binning.names:
                return self.unroll_binning(key, binning)
        if key not in self.keys:
            if key in self.all_keys:
                self.auto_translate(key)
            else:
                if key in self._aux_data:
                    return self._aux_data[key]
                raise KeyError('Key "%s" not present in Container "%s"' % (key, self.name))
        if not self.validity[key].get(hash(self.representation), False):
            self.auto_translate(key)
        return self.current_data[key]

    def unroll_binning(self, key, binning):
        """Unrolled (row-major flattened) bin centres of dimension `key` (container.py:769-773)."""
        ck = ("unroll", hash(binning), key)
        if ck not in self._index_cache:
            grid = binning.meshgrid(entity="weighted_centers", attach_units=False)
            self._index_cache[ck] = ((), self._to_device(grid[binning.index(key)].ravel()))
        return self._index_cache[ck][1]

    def get_hist(self, key):
        """(host ndarray reshaped to the binning's shape, binning) (container.py:776-790)."""
        assert self.is_map, "Cannot retrieve hists from non-map data"
        binning = self.representation
        data = self[key].detach().cpu().numpy()
        full_shape = list(binning.shape) + ([-1] if data.ndim > 1 else [])
        return data.reshape(full_shape), binning

    def get_map(self, key, error=None):
        hist, binning = self.get_hist(key)
        error_hist = np.abs(self.get_hist(error)[0]) if error is not None else None
        assert hist.ndim == binning.num_dims
        return Map(name=self.name, hist=hist, error_hist=error_hist, binning=binning)

    def __iter__(self):
        return iter(self.keys)

    # --------------------------------------------------------------------- translation ---
    def bin_index(self, binning, policy="generic"):
        """Flat row-major bin index (int32, -1 outside) of every event on `binning`, cached.
        Equivalent to what the reference recomputes inside every histogram()/lookup() call; ``policy`` selects the
        edge rule (see ``regularized_dims``): "generic" for the container translations, "hist" for ``utils.hist``."""
        from pisa_b200 import ops
        ck = ("index", hash(binning), policy)
        hit = self._index_cache.get(ck)
        if hit is not None:
            return hit[1]
        saved = self.representation
        self.representation = "events"
        coords = [self[name] for name in binning.names]
        self.representation = saved
        b, _keep = ops.make_binning(regularized_dims(binning, policy), coords[0].device)
        idx = ops.hist_index(b, coords)
        self._index_cache[ck] = (tuple(binning.names), idx)
        self._index_cache_names.update(binning.names)
        return idx

    def bin_plan(self, binning, policy="generic"):
        """``ops.HistPlan`` of ``bin_index(binning, policy)`` (cached with it; None when the binning is not plannable):
        the fit-loop form of the histogram, see ``pisab_hist_plan_build``."""
        from pisa_b200 import ops
        ck = ("plan", hash(binning), policy)
        hit = self._index_cache.get(ck)
        if hit is not None:
            return hit[1]
        plan = ops.hist_plan(self.bin_index(binning, policy), binning.size)
        self._index_cache[ck] = (tuple(binning.names), plan)
        return plan

    def translate(self, key, src_representation):
        assert hash(src_representation) in self.representation_keys
        dest_representation = self.representation
        if hash(src_representation) == hash(dest_representation):
            return
        from_map = isinstance(src_representation, MultiDimBinning)
        to_map = isinstance(dest_representation, MultiDimBinning)
        mode = self.translation_modes[key]
        if mode == "average":
            if from_map and to_map:
                out = self.resample(key, src_representation, dest_representation)
            elif to_map:
                out = self.array_to_binned(key, src_representation, dest_representation)
            elif from_map:
                out = self.binned_to_array(key, src_representation, dest_representation)
            elif src_representation == "events" and dest_representation == "log_events":
                self.representation = "events"
                out = torch.log(self[key])
            elif src_representation == "log_events" and dest_representation == "events":
                self.representation = "log_events"
                out = torch.exp(self[key])
            else:
                raise NotImplementedError("Translating %s to %s in 'average' mode!"
                                          % (src_representation, dest_representation))
        elif mode == "sum":
            if from_map and to_map:
                raise NotImplementedError("Map to Map in sum mode needs to integrate over bins.")
            if to_map:
                out = self.array_to_binned(key, src_representation, dest_representation, averaged=False)
            else:
                raise NotImplementedError("Translating %s to %s in 'sum' mode!"
                                          % (src_representation, dest_representation))
        else:
            raise ValueError("Unknown translation mode for variable '%s': '%s'!" % (key, mode))
        self.representation = dest_representation
        self[key] = out
        self.validity[key][hash(src_representation)] = True

    def auto_translate(self, key):
        src = self.find_valid_representation(key)
        if src is None:
            raise Exception("No valid representation for %s in container" % key)
        self.translate(key, src)

    def find_valid_representation(self, key):
        best, representation = np.inf, None
        for h, ok in self.validity[key].items():
            if ok and self.precedence[h] < best:
                best, representation = self.precedence[h], self._representations[h]
        return representation

    def resample(self, key, src_representation, dest_representation):
        """binned -> binned (container.py:913-931 -> translation.resample, translation.py:49-85): the old bin
        centres are histogrammed into the new binning (mean of the values where more than one old bin lands in a
        new bin); every other new bin takes the value of the old bin its own centre falls into."""
        from pisa_b200 import ops
        if src_representation.names != dest_representation.names:
            raise ValueError("cannot translate betwen %s and %s" % (src_representation, dest_representation))
        self.representation = src_representation
        old_sample = [self[name] for name in src_representation.names]
        weights = self[key]
        self.representation = dest_representation
        new_sample = [self[name] for name in dest_representation.names]
        if weights.dim() != 1:
            raise NotImplementedError("resampling of vector-valued maps")
        dev = weights.device
        # (both keep-alive lists must outlive the hist_index calls below: the structs hold raw edge pointers)
        new_b, _keep_new = ops.make_binning(regularized_dims(dest_representation, "generic"), dev)
        old_b, _keep_old = ops.make_binning(regularized_dims(src_representation, "generic"), dev)
        into_new = ops.hist_index(new_b, old_sample)
        n_bins = dest_representation.size
        summed, _ = ops.hist_accumulate(into_new, weights, n_bins, want_w2=False)
        counts, _ = ops.hist_accumulate(into_new, None, n_bins, want_w2=False)
        mean = torch.nan_to_num(summed / counts, nan=0.0, posinf=0.0, neginf=0.0)
        looked_up = ops.lookup(ops.hist_index(old_b, new_sample), weights)
        return torch.where(counts > 1, mean.to(looked_up.dtype), looked_up).to(TDTYPE)

    def array_to_binned(self, key, src_representation, dest_representation, averaged=True):
        """events -> binned: weighted histogram, divided by the counts when `averaged`
        (translation.histogram, translation.py:90-129)."""
        from pisa_b200 import ops
        assert src_representation in self.array_representations
        assert isinstance(dest_representation, MultiDimBinning)
        idx = self.bin_index(dest_representation)
        plan = self.bin_plan(dest_representation)       # bin-sorted tiles + static counts (None above 256 bins)
        self.representation = "events"
        weights = self[key]
        n_bins = dest_representation.size
        cols = [weights] if weights.dim() == 1 else [weights[:, i].contiguous() for i in range(weights.shape[1])]
        hists = [ops.hist_accumulate(idx, w, n_bins, want_w2=False, plan=plan)[0] for w in cols]
        if averaged:
            counts, _ = ops.hist_accumulate(idx, None, n_bins, want_w2=False, plan=plan)
            # flat_hist / counts with nan_to_num (translation.py:118-127)
            hists = [torch.nan_to_num(h / counts, nan=0.0, posinf=0.0, neginf=0.0) for h in hists]
        out = hists[0] if weights.dim() == 1 else torch.stack(hists, dim=1)
        self.representation = dest_representation
        return out.to(TDTYPE)

    def binned_to_array(self, key, src_representation, dest_representation):
        """binned -> events: value of the bin each event falls into, 0 outside
        (translation.lookup, translation.py:228-344)."""
        from pisa_b200 import ops
        self.representation = src_representation
        flat_hist = self[key]
        idx = self.bin_index(src_representation)
        self.representation = dest_representation
        return ops.lookup(idx, flat_hist)


class VirtualContainer:
    """Linked containers behave like one (container.py:363-448): reads come from the first one,
    writes go to all."""

    def __init__(self, name, containers):
        self.name = name
        for c in containers:
            if c.linked:
                raise ValueError("Cannot link container %s since it is already linked" % c.name)
            c.linked = True
        self.containers = containers

    def __repr__(self):
        return "VirtualContainer containing %s" % [c.name for c in self]

    def unlink(self):
        for c in self:
            c.linked = False

    def __iter__(self):
        return iter(self.containers)

    def __getitem__(self, key):
        return self.containers[0][key]

    def __setitem__(self, key, value):
        for c in self:
            c[key] = value

    def set_aux_data(self, key, val):
        for c in self:
            c.set_aux_data(key, val)

    def mark_changed(self, key):
        # every linked container gets its OWN copy (np.copy in the reference, container.py:427-433): stages that
        # later mutate one container's array in place (aeff, prob3.apply) must not reach the others
        src = self.containers[0][key]
        for c in self.containers[1:]:
            c[key] = src.clone()
        for c in self:
            c.mark_changed(key)

    def mark_valid(self, key):
        for c in self:
            c.mark_valid(key)

    @property
    def representation(self):
        return self.containers[0].representation

    @representation.setter
    def representation(self, representation):
        for c in self:
            c.representation = representation

    shape = property(lambda self: self.containers[0].shape)
    size = property(lambda self: int(np.prod(self.shape)))
    is_map = property(lambda self: self.containers[0].is_map)

    def bin_index(self, binning, policy="generic"):
        return self.containers[0].bin_index(binning, policy)

    def bin_plan(self, binning, policy="generic"):
        return self.containers[0].bin_plan(binning, policy)


class ContainerSet:
    def __init__(self, name, containers=None, representation=None):
        self.name = name
        self.linked_containers = []
        self.containers = []
        for c in (containers or []):
            self.add_container(c)
        self._representation = None
        self.representation = representation
        self._glob_aux_data = {}

    def __repr__(self):
        return "ContainerSet containing %s" % [c.name for c in self]

    @property
    def is_map(self):
        if len(self.containers):
            return self.containers[0].is_map
        return None

    def add_container(self, container):
        if container.name in self.names:
            raise ValueError("container with name %s already exists" % container.name)
        self.containers.append(container)

    @property
    def representation(self):
        return self._representation

    @representation.setter
    def representation(self, representation):
        self._representation = representation
        if representation is None:
            return
        for c in self:
            c.representation = representation

    names = property(lambda self: [c.name for c in self.containers])

    def get_shared_keys(self, rep_indep=True):
        if len(self.containers) == 0:
            return ()
        return tuple(set.intersection(*[
            set(c.all_keys_incl_aux_data if rep_indep else c.keys_incl_aux_data) for c in self.containers]))

    def link_containers(self, key, names):
        link_names = [n for n in names if n in self.names]
        containers = [self[n] for n in link_names]
        if containers:
            self.linked_containers.append(VirtualContainer(key, containers))

    def unlink_containers(self):
        for c in self.linked_containers:
            c.unlink()
        self.linked_containers = []

    def __getitem__(self, key):
        if key in self.names:
            return self.containers[self.names.index(key)]
        for c in self.linked_containers:
            if c.name == key:
                return c
        if key in self._glob_aux_data:
            return self._glob_aux_data[key]
        raise KeyError("No name `%s` in container" % key)

    def __setitem__(self, key, data):
        if key in self.names:
            raise KeyError("`%s` is a container name. If you want to update a container use self.containers." % key)
        if key in [c.name for c in self.linked_containers]:
            raise KeyError("`%s` is a linked container name and can't be overwritten." % key)
        self._glob_aux_data[key] = data

    def __iter__(self):
        return iter([c for c in self.containers if not c.linked] + self.linked_containers)

    def get_mapset(self, key, error=None):
        """MapSet with one Map per container (container.py:339-355).  The per-container histograms are
        gathered on the device and read back with ONE device->host copy per key (a copy per container
        costs a stream synchronisation each: 24 per template for 12 containers with errors)."""
        conts = list(self)
        if not conts or not all(c.is_map for c in conts):
            return MapSet(name=self.name, maps=[c.get_map(key, error=error) for c in conts])

        def fetch(k):
            tensors = [c[k].detach() for c in conts]
            if len({(tuple(t.shape), t.dtype) for t in tensors}) != 1:
                return [t.cpu().numpy() for t in tensors]
            return list(torch.stack(tensors).cpu().numpy())

        hists = fetch(key)
        errs = fetch(error) if error is not None else [None] * len(conts)
        maps = []
        for c, h, e in zip(conts, hists, errs):
            binning = c.representation
            full_shape = list(binning.shape) + ([-1] if h.ndim > 1 else [])
            h = h.reshape(full_shape)
            assert h.ndim == binning.num_dims
            maps.append(Map(name=c.name, hist=h, error_hist=None if e is None else np.abs(e.reshape(full_shape)),
                            binning=binning))
        return MapSet(name=self.name, maps=maps)

    glob_aux_data_keys = property(lambda self: self._glob_aux_data.keys())
