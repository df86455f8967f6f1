"""Stand-alone histogram kernels at 1e8 events (CUDA events, best of 5): replicated-slot kernel, planned kernel
(bin-sorted tiles), exact fixed-point kernel (3200 bins), plus plan build time."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pisa_b200 import ops

def timeit(f, reps=5):
    for _ in range(2): f()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)

dev = torch.device("cuda:0")
n = 100_000_000
g = torch.Generator(device=dev); g.manual_seed(1)
for dtype, wb in ((torch.float64, 8), (torch.float32, 4)):
    w = torch.rand(n, generator=g, device=dev, dtype=torch.float64).to(dtype)
    for n_bins in (128, 256, 512, 1024):
        idx = torch.randint(-1, n_bins, (n,), generator=g, device=dev, dtype=torch.int32)
        t_slot = timeit(lambda: ops.hist_accumulate(idx, w, n_bins))
        t_plan_build = timeit(lambda: ops.hist_plan(idx, n_bins), reps=2)
        plan = ops.hist_plan(idx, n_bins)
        t_plan = timeit(lambda: ops.hist_accumulate(idx, w, n_bins, plan=plan))
        a, a2 = ops.hist_accumulate(idx, w, n_bins); b, b2 = ops.hist_accumulate(idx, w, n_bins, plan=plan)
        rel = float(((a - b).abs() / a.abs()).max())
        alg = (4 + wb) * n
        print("%s %4d bins: slots %.3f ms (%.0f GB/s of %d B/event) | planned %.3f ms (%.0f GB/s algorithmic, %.0f GB/s moved at %d B/event) %.2e ev/s | plan build %.1f ms | max rel diff %.1e"
              % (str(dtype)[6:], n_bins, t_slot, alg / t_slot / 1e6, 4 + wb, t_plan, alg / t_plan / 1e6, (2 + wb) * n / t_plan / 1e6, 2 + wb, n / t_plan * 1e3, t_plan_build, rel), flush=True)
    idx = torch.randint(-1, 3200, (n,), generator=g, device=dev, dtype=torch.int32)
    t_fix = timeit(lambda: ops.hist_accumulate(idx, w, 3200))
    print("%s 3200 bins exact fixed-point: %.3f ms (%.0f GB/s of %d B/event incl. the max|w| pass)" % (str(dtype)[6:], t_fix, (4 + wb) * n / t_fix / 1e6, 4 + wb), flush=True)
    t_build = timeit(lambda: ops.hist_plan(idx, 3200), reps=2)
    plan = ops.hist_plan(idx, 3200)
    t_sorted = timeit(lambda: ops.hist_accumulate(idx, w, 3200, plan=plan))
    a, a2 = ops.hist_accumulate(idx, w, 3200); b, b2 = ops.hist_accumulate(idx, w, 3200, plan=plan)
    print("%s 3200 bins sorted plan (exact, one atomic pair per warp): %.3f ms (%.0f GB/s of %d B/event algorithmic) | plan build %.1f ms | bit-identical to the unplanned exact sums: %s"
          % (str(dtype)[6:], t_sorted, (4 + wb) * n / t_sorted / 1e6, 4 + wb, t_build, bool(torch.equal(a, b) and torch.equal(a2, b2))), flush=True)
