"""N-GPU check + latency of the one-launch peer-memory exchange (torchrun --nproc-per-node N scratch/check_exchange.py)."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pisa_b200 import distributed as D

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
ok = True
g = torch.Generator(device=dev); g.manual_seed(100 + rank)
for it, count in enumerate([3072, 1, 17, 3072, 65536, 307200, 3072, 3072]):
    buf = torch.rand(count, generator=g, device=dev, dtype=torch.float64) * 10.0 ** float(it - 3)
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf)
    want = gathered[0].clone()
    for r in range(1, world):
        want += gathered[r]
    got = D.combine_histograms(buf.clone())
    same = torch.equal(got, want)
    ref0 = got.clone(); dist.broadcast(ref0, 0)
    ident = torch.equal(got, ref0)
    ok = ok and same and ident
    if rank == 0:
        print("count %7d: equals rank-ordered sum %s, identical on all ranks %s, mode %s" % (count, same, ident, D.exchange_mode()), flush=True)
# back-to-back calls (epoch / parity logic) with different data each time
buf = torch.zeros(3072, device=dev, dtype=torch.float64)
acc = torch.zeros_like(buf)
for it in range(2000):
    buf.fill_(float(rank + 1) * (it + 1))
    D.combine_histograms(buf)
    acc += buf
torch.cuda.synchronize()
expect = sum(r + 1 for r in range(world)) * sum(i + 1 for i in range(2000))
ok = ok and bool((acc == expect).all())
if rank == 0:
    print("2000 back-to-back exchanges:", bool((acc == expect).all()), "status", D.exchange_status(), flush=True)

def lat(mode, count, reps=500):
    os.environ["PISAB_EXCHANGE"] = mode
    b = torch.ones(count, device=dev, dtype=torch.float64)
    for _ in range(20): D.combine_histograms(b)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(reps): D.combine_histograms(b)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3, (time.perf_counter() - t0) / reps * 1e6
for count in (3072, 307200):
    p = lat("peer", count); n = lat("nccl", count)
    if rank == 0:
        print("count %7d on %d GPUs: peer kernel %.1f us (host %.1f us) | all_gather + sum kernel %.1f us (host %.1f us)" % (count, world, p[0], p[1], n[0], n[1]), flush=True)
os.environ["PISAB_EXCHANGE"] = "peer"
if rank == 0:
    print("EXCHANGE_OK" if ok else "EXCHANGE_FAILED", flush=True)
dist.destroy_process_group()
