"""Multi-GPU plumbing of the hot path: one process per GPU, events sharded, ONE exchange.

Events are independent, so every rank owns a fixed contiguous slice of every flavour container
and evaluates it with the same kernels; per-launch constants are replicated.  The only data that
crosses NVLink is the ``[containers, 2, n_bins]`` float64 histogram buffer (24 KB for the
``dragon_datarelease`` binning), combined once per template.

Two combination modes:
  * ``deterministic=True`` (default): ``all_gather`` the per-rank buffers and sum them in rank
    order on every rank -- the N-GPU histogram is then bit-reproducible run to run and identical on
    all ranks, whatever algorithm / channel count NCCL picks;
  * ``deterministic=False``: a plain ``all_reduce(SUM)`` (NVLS in-switch reduction when available).
Both are latency-bound at this size (tens of microseconds against a >= 18 ms step).

The same code runs over ``gloo`` on CPU tensors (tests/test_multi_rank.py).
"""
import os

import torch
import torch.distributed as dist

__all__ = ["init_from_env", "world", "shard_slice", "shard_arrays", "combine_histograms", "enable_event_sharding",
           "event_sharding", "local_slice"]

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


def combine_histograms(buf, deterministic=True):
    """Sum the per-rank histogram buffers in place (the single exchange step of the path)."""
    rank, world_size = world()
    if world_size == 1:
        return buf
    if not deterministic:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        return buf
    gathered = torch.empty((world_size,) + tuple(buf.shape), dtype=buf.dtype, device=buf.device)
    dist.all_gather(list(gathered.unbind(0)), buf.contiguous())
    acc = gathered[0].clone()
    for r in range(1, world_size):   # fixed rank order: bit-identical on every rank and every run
        acc += gathered[r]
    buf.copy_(acc)
    return buf
