#!/bin/bash
# round 2, GPU call 22: planned histogram up to 1024 bins, bench variants with parity checks
mkdir -p gpurun_out
O=gpurun_out/r02_run22.txt
{
echo "== pytest hist"; timeout 900 python -m pytest tests/test_gpu_hist.py tests/test_gpu_pipeline.py -m gpu -q 2>&1 | tail -5
echo "== bench hist"; timeout 900 python scratch/bench_hist.py
echo "== bench"; timeout 1500 python bench.py --no-cpu-baseline > gpurun_out/r02_bench22.json 2> gpurun_out/r02_bench22.err; echo "rc=$?"; tail -3 gpurun_out/r02_bench22.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/r02_bench22.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "parity", d.get("parity_check"))
for k, v in (d.get("variants") or {}).items():
    print(k, {kk: v[kk] for kk in v if kk in ("value", "ms_per_step", "parity_check")})
P
} > $O 2>&1
tail -40 $O
