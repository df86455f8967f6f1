"""Concurrent H2D bandwidth of all ranks with and without binding each rank to its GPU's CPU affinity.
torchrun --nproc-per-node N scratch/h2d_numa_probe.py"""
import os, time
import torch, torch.distributed as dist
import pynvml
rank = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(rank)
ncpu = os.cpu_count()
words = (ncpu + 63) // 64
mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
cpus = [i for i in range(ncpu) if (mask[i // 64] >> (i % 64)) & 1]
try:
    numa = pynvml.nvmlDeviceGetNumaNodeId(h)
except Exception:
    numa = -1
def measure(tag):
    n = 1 << 30
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    host.fill_(1)                       # first touch on the current CPU set
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    dev.copy_(host, non_blocking=True); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(4): dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    gbs = 4 * n / dt / 1e9
    out = [None] * world
    if world > 1: dist.all_gather_object(out, (rank, gbs))
    else: out = [(rank, gbs)]
    if rank == 0:
        print(tag, " ".join("%.1f" % g for _, g in sorted(out)), "sum %.1f GB/s" % sum(g for _, g in out), flush=True)
if rank == 0: print("cpus", ncpu, flush=True)
print("rank", rank, "numa", numa, "affinity", "%d-%d (%d cpus)" % (cpus[0], cpus[-1], len(cpus)) if cpus else "none",
      "current", len(os.sched_getaffinity(0)), flush=True)
measure("unbound")
if cpus:
    os.sched_setaffinity(0, cpus)
measure("bound  ")
