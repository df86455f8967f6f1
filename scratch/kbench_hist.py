"""Variants of the histogram kernel (compile flags) timed at 1e8 events."""
import json, os, subprocess, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scratch"))
from pisa_b200 import _lib, build as B
from kbench import build_variant
from kbench_stage import timeit
variants = json.loads(sys.argv[1])
n = 100_000_000
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)
idx = torch.randint(-1, 128, (n,), generator=g, device=dev, dtype=torch.int32)
w = torch.rand(n, generator=g, device=dev, dtype=torch.float64)
ref = None
for tag, flags in variants.items():
    try:
        path = build_variant(tag, flags)
        _lib._lib = None; B.LIB = path; _lib._build.LIB = path
        from pisa_b200 import ops
        ops._workspaces.clear()
        t = timeit(lambda: ops.hist_accumulate(idx, w, 128))
        h, h2 = ops.hist_accumulate(idx, w, 128)
        chk = float(h.sum())
        print("%-24s %8.3f ms  %7.1f GB/s  checksum %.12e" % (tag, t, 12 * n / t / 1e6, chk), flush=True)
    except Exception as e:
        print(tag, "FAILED", repr(e)[:200], flush=True)
