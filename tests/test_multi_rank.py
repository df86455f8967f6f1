"""N > 1 host logic on CPU: world_size-2 ``gloo`` process group.

The data path of a rank is the CUDA library and cannot run here; what CAN be covered without a GPU
is everything that differs between 1 and N ranks: the fixed shard boundaries, the single exchange
step (``combine_histograms``, deterministic and plain modes) and the invariant the multi-GPU path
relies on -- the sum over shards of per-shard histograms equals the unsharded histogram.  The
per-shard histograms are produced by the CPU oracle (checker used as stand-in for the kernel).
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pisa_b200.distributed import combine_histograms, init_from_env, shard_arrays, shard_slice, world  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_template(events_by_container):
    """[containers, 2, 128] histogram of the reference chain (layers -> prob3 -> reweight -> hist)."""
    import oracle
    from pisa_b200.utils import synthetic as syn
    prem = np.loadtxt(os.path.join(ROOT, "pisa_b200", "resources", "osc", "PREM_12layer.dat"))
    L = oracle.OracleLayers(prem, syn.EARTH["detector_depth"], syn.EARTH["prop_height"])
    L.setElecFrac(syn.EARTH["YeI"], syn.EARTH["YeO"], syn.EARTH["YeM"])
    dm, mix, mat_pot = syn.osc_matrices(nsi=syn.STD_NSI)
    zc, zf = np.zeros((3, 3), dtype=np.complex128), np.zeros((3, 3))
    out = np.zeros((len(events_by_container), 2, 128))
    for c, (nubar, flav, ev) in enumerate(events_by_container):
        if ev["true_energy"].shape[0] == 0:
            continue
        _, den, dis = L.calcLayers(ev["true_coszen"])
        prob = oracle.propagate_array(dm, mix, mat_pot, -1, zc, zf, nubar, ev["true_energy"], den, dis, n_threads=1)
        w = ev["weights"] * (ev["nu_flux"][:, 0] * prob[:, 0, flav] + ev["nu_flux"][:, 1] * prob[:, 1, flav])
        ie = oracle.digitize_irregular(ev["reco_energy"], syn.DRAGON_E_EDGES)
        i2, _ = oracle.regular_index([ev["reco_coszen"], ev["pid"]], [-1.0, -0.5], [1.0, 1.5], [8, 2])
        idx = np.where((ie >= 0) & (ie < 8) & (i2 >= 0), ie * 16 + i2, -1)
        out[c, 0] = oracle.accumulate(idx, w, 128)
        out[c, 1] = oracle.accumulate(idx, w * w, 128)
    return out


def _containers(n):
    from pisa_b200.utils import synthetic as syn
    return [(nubar, flav, syn.make_events_numpy(n + 7 * c, seed=100 + c))
            for c, (name, nubar, flav) in enumerate(syn.CONTAINERS[:3] + syn.CONTAINERS[6:8])]


def _worker(rank, world_size, port, tmpdir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world_size), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(1)
    r, w = init_from_env(backend="gloo")
    assert (r, w) == (rank, world_size) == world()
    full = _containers(501)
    mine = [(nubar, flav, shard_arrays(ev, rank, world_size)) for nubar, flav, ev in full]
    local = torch.from_numpy(_oracle_template(mine))
    det = combine_histograms(local.clone(), deterministic=True)
    plain = combine_histograms(local.clone(), deterministic=False)
    # a second deterministic combination must be bit-identical (fixed rank order)
    det2 = combine_histograms(local.clone(), deterministic=True)
    assert torch.equal(det, det2)
    np.savez(os.path.join(tmpdir, "rank%d.npz" % rank), det=det.numpy(), plain=plain.numpy(), local=local.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_slice_partitions_every_length():
    for n in (0, 1, 2, 7, 12, 1000, 99999996):
        for w in (1, 2, 3, 4, 8):
            edges = [shard_slice(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 0
    with pytest.raises(ValueError):
        shard_slice(10, 2, 2)
    with pytest.raises(ValueError):
        shard_arrays(dict(a=np.zeros(3), b=np.zeros(4)), 0, 2)


def test_single_rank_is_a_no_op():
    buf = torch.arange(12, dtype=torch.float64).reshape(2, 2, 3)
    assert combine_histograms(buf.clone()).equal(buf)
    assert world() == (0, 1)


@pytest.mark.timeout(300)
def test_two_rank_gloo_sharded_template_equals_unsharded(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in (0, 1))
    # every rank ends with the same buffer, bit for bit, in deterministic mode
    assert np.array_equal(r0["det"], r1["det"])
    # deterministic mode == rank-ordered sum of the local buffers
    assert np.array_equal(r0["det"], r0["local"] + r1["local"])
    # plain all_reduce agrees to rounding
    np.testing.assert_allclose(r0["plain"], r0["det"], rtol=1e-15, atol=0)
    # and the sharded template equals the unsharded one (different summation order: 1e-13, inside 1e-10)
    whole = _oracle_template(_containers(501))
    assert whole[:, 0].sum() > 0
    np.testing.assert_allclose(r0["det"], whole, rtol=1e-12, atol=1e-300)


def test_event_sharding_switch_and_local_slice():
    """Stage-API sharding is opt-in: without a process group (or switched off) every rank keeps all events."""
    from pisa_b200 import distributed as D
    D.enable_event_sharding(True)
    try:
        assert not D.event_sharding()                     # one rank: nothing to shard
        assert D.local_slice(10) == slice(0, 10)
    finally:
        D.enable_event_sharding(False)
    assert D.local_slice(7) == slice(0, 7)
