#!/bin/bash
# round 2, GPU call 13: hypersurface pin, sorted large-binning plan, warp-aggregated exact adds in the fused kernel
mkdir -p gpurun_out
O=gpurun_out/r02_run13.txt
{
echo "== pytest (all)"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15
echo "== bench hist"; timeout 900 python scratch/bench_hist.py
echo "== bench variants"; timeout 900 python bench.py --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r02_bench13.json 2>gpurun_out/r02_bench13.err; echo rc=$?
python - <<'P'
import json
d = json.loads(open("gpurun_out/r02_bench13.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"])
for k, v in (d.get("variants") or {}).items():
    print(k, {kk: v[kk] for kk in ("value", "ms_per_step") if kk in v})
P
} > $O 2>&1
tail -60 $O
