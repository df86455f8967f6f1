"""One launch of the round-2 histogram kernels at 1e8 events for a single `ncu --set full` capture: planned histogram
(f64 / f32 weights), exact fixed-point accumulation (3200 bins), the scale_weights (aeff) and joint-index kernels."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pisa_b200 import ops
dev = torch.device("cuda:0")
n = 100_000_000
g = torch.Generator(device=dev); g.manual_seed(1)
idx = torch.randint(-1, 128, (n,), generator=g, device=dev, dtype=torch.int32)
idx_big = torch.randint(-1, 3200, (n,), generator=g, device=dev, dtype=torch.int32)
w = torch.rand(n, generator=g, device=dev, dtype=torch.float64)
w32 = w.float()
a = torch.rand(n, generator=g, device=dev, dtype=torch.float64)
plan = ops.hist_plan(idx, 128)
def all_ops():
    ops.hist_accumulate(idx, w, 128, plan=plan)
    ops.hist_accumulate(idx, w32, 128, plan=plan)
    ops.hist_accumulate(idx_big, w, 3200)
    ops.scale_weights(w, a, 1.0)
    ops.hist_plan(idx, 128)
all_ops(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
all_ops(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
