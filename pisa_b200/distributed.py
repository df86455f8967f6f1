"""Multi-GPU plumbing of the hot path: one process per GPU, events sharded, ONE exchange.

Events are independent, so every rank owns a fixed contiguous slice of every flavour container
and evaluates it with the same kernels; per-launch constants are replicated.  The only data that
crosses NVLink is the ``[containers, 2, n_bins]`` float64 histogram buffer (24 KB for the
``dragon_datarelease`` binning), combined once per template.

Combination modes (``combine_histograms``):
  * ``deterministic=True`` (default), CUDA buffers: ONE kernel launch per rank over NVLink peer memory
    (``csrc/exchange.cu``: every rank stores its buffer into a slot of every peer's exchange buffer, publishes a
    system-scope flag, waits for the peers' flags and sums the slots in rank order) -- bit-reproducible run to run and
    identical on all ranks.  The exchange buffers are mapped once through CUDA IPC handles gathered with
    ``torch.distributed`` (plumbing).  ``PISAB_EXCHANGE=nccl`` (or a failed peer mapping) selects the library form:
    ``all_gather`` + one rank-ordered sum kernel (2 launches);
  * CPU tensors (``gloo``, the host-logic tests): ``all_gather`` + rank-ordered sum in torch;
  * ``deterministic=False``: a plain ``all_reduce(SUM)`` (NVLS in-switch reduction when available).
All are latency-bound at this size (tens of microseconds against an 11 ms step at 1e8 events; the one-launch form
matters for analysis-size templates of ~45 us).

The same code runs over ``gloo`` on CPU tensors (tests/test_multi_rank.py).
"""
import os

import torch
import torch.distributed as dist

__all__ = ["init_from_env", "world", "shard_slice", "shard_arrays", "combine_histograms", "enable_event_sharding",
           "event_sharding", "local_slice", "PeerExchange", "exchange_mode", "exchange_status"]

# Stage API: with sharding on, the event-mode loaders keep only this rank's slice of every container and
# ``utils.hist`` (or the fused engine) exchanges the binned results once per template.  Off by default, so that a
# process group initialised for another purpose does not change what a Pipeline computes.
_SHARD_EVENTS = os.environ.get("PISAB_SHARD_EVENTS", "0").strip().lower() not in ("", "0", "false", "no")


def init_from_env(backend=None, device=None):
    """Initialise ``torch.distributed`` from RANK / WORLD_SIZE / MASTER_* (torchrun); no-op for one rank.
    Returns (rank, world_size)."""
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world_size > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl" and device is not None:
            kwargs["device_id"] = device
        dist.init_process_group(backend, rank=rank, world_size=world_size, **kwargs)
    return rank, world_size


def world():
    """(rank, world_size) of the initialised process group, (0, 1) otherwise."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def enable_event_sharding(on=True):
    """Switch the sharding of events over the ranks of the process group on or off (also: PISAB_SHARD_EVENTS=1)."""
    global _SHARD_EVENTS
    _SHARD_EVENTS = bool(on)


def event_sharding():
    """True when events are sharded: switched on AND more than one rank."""
    return _SHARD_EVENTS and world()[1] > 1


def local_slice(n):
    """``slice`` of the n events of a container that this rank keeps (all of them without sharding)."""
    if not event_sharding():
        return slice(0, int(n))
    rank, world_size = world()
    start, stop = shard_slice(n, rank, world_size)
    return slice(start, stop)


def shard_slice(n, rank, world_size):
    """Contiguous slice [start, stop) of n events owned by ``rank``.  Boundaries depend on
    (n, world_size) only, so a re-run shards identically (fixed summation order per rank)."""
    n, rank, world_size = int(n), int(rank), int(world_size)
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside [0, %d)" % (rank, world_size))
    base, rem = divmod(n, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_arrays(arrays, rank, world_size):
    """Slice every per-event array of one container (dict of arrays whose first axis is the event
    axis) to this rank's share."""
    n = None
    for a in arrays.values():
        if n is None:
            n = a.shape[0]
        elif a.shape[0] != n:
            raise ValueError("per-event arrays of one container must have the same length")
    start, stop = shard_slice(n or 0, rank, world_size)
    return {k: a[start:stop] for k, a in arrays.items()}


class PeerExchange:
    """This rank's end of the peer-memory exchange (``pisab_exchange_*``): buffer, IPC handshake, one-launch
    rank-ordered all-reduce of float64 CUDA buffers up to ``capacity`` values."""

    def __init__(self, device, capacity):
        import ctypes
        from . import _lib
        self._lib, self._ct = _lib, ctypes
        rank, world_size = world()
        self.capacity, self.device = int(capacity), device
        self.ctx = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        with torch.cuda.device(device):
            _lib.check(_lib.load().pisab_exchange_create(rank, world_size, self.capacity, ctypes.byref(self.ctx), handle))
            handles = [None] * world_size
            dist.all_gather_object(handles, bytes(handle))
            blob = b"".join(handles)
            rc = _lib.load().pisab_exchange_connect(self.ctx, blob)
        ok = torch.tensor([1 if rc == 0 else 0], device=device if dist.get_backend() == "nccl" else "cpu")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)       # all ranks take the same path
        if int(ok) == 0:
            self.close(collective=False)
            raise RuntimeError("peer mapping of the exchange buffers failed on at least one rank: %s"
                               % (_lib.load().pisab_last_error().decode() if rc else "another rank"))

    def allreduce(self, buf):
        from . import ops
        self._lib.check(self._lib.load().pisab_exchange_allreduce(self.ctx, self._ct.c_void_p(buf.data_ptr()),
                                                                   buf.numel(), ops._stream()))
        return buf

    def status(self):
        return int(self._lib.load().pisab_exchange_status(self.ctx))

    def close(self, collective=True):
        """Orderly shutdown: every rank unmaps its peers, the ranks synchronise, then every rank frees its own buffer
        (an exported buffer must not be freed while a peer still maps it).  Collective unless ``collective=False``."""
        if self.ctx:
            self._lib.load().pisab_exchange_disconnect(self.ctx)
            if collective and dist.is_initialized():
                dist.barrier()
            self._lib.load().pisab_exchange_destroy(self.ctx)
            self.ctx = None


_peer = {"exchange": None, "failed": False}


def exchange_mode():
    """How CUDA histogram buffers are combined: "peer" (one kernel over NVLink peer memory) or "nccl"."""
    if os.environ.get("PISAB_EXCHANGE", "peer").strip().lower() == "nccl" or _peer["failed"]:
        return "nccl"
    return "peer"


def _peer_exchange(device, count):
    ex = _peer["exchange"]
    if ex is not None and ex.capacity >= count and ex.device == device:
        return ex
    if ex is not None:
        ex.close()
    try:
        _peer["exchange"] = PeerExchange(device, max(int(count), 1 << 19))   # 4 MB per slot: a 100-template scan fits
    except RuntimeError as exc:
        import warnings
        warnings.warn("pisa_b200: peer-memory exchange unavailable (%s); using all_gather + sum kernel" % exc)
        _peer["exchange"], _peer["failed"] = None, True
    return _peer["exchange"]


def exchange_status():
    """0 = every peer wait so far succeeded (None when the peer exchange is not in use).  Synchronises the device."""
    ex = _peer["exchange"]
    return None if ex is None else ex.status()


def combine_histograms(buf, deterministic=True):
    """Sum the per-rank histogram buffers in place (the single exchange step of the path)."""
    rank, world_size = world()
    if world_size == 1:
        return buf
    if not deterministic:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        return buf
    if buf.is_cuda and buf.dtype == torch.float64 and buf.is_contiguous():
        if exchange_mode() == "peer":
            ex = _peer_exchange(buf.device, buf.numel())
            if ex is not None:
                return ex.allreduce(buf)
        from . import _lib, ops
        import ctypes
        gathered = torch.empty((world_size,) + tuple(buf.shape), dtype=buf.dtype, device=buf.device)
        dist.all_gather_into_tensor(gathered, buf)
        _lib.check(_lib.load().pisab_sum_slots(ctypes.c_void_p(gathered.data_ptr()), world_size, buf.numel(),
                                               ctypes.c_void_p(buf.data_ptr()), ops._stream()))
        return buf
    gathered = torch.empty((world_size,) + tuple(buf.shape), dtype=buf.dtype, device=buf.device)
    dist.all_gather(list(gathered.unbind(0)), buf.contiguous())
    acc = gathered[0].clone()
    for r in range(1, world_size):   # fixed rank order: bit-identical on every rank and every run
        acc += gathered[r]
    buf.copy_(acc)
    return buf
