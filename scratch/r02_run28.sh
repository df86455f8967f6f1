#!/bin/bash
# round 2, GPU call 28: WarpHist::add2 in the two-events-per-thread kernel
mkdir -p gpurun_out
O=gpurun_out/r02_run28.txt
{
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
for i in 1 2; do
timeout 600 python bench.py --dtype f32 --no-cpu-baseline --no-e2e --no-variants --no-parity 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('f32', d['value'], d['ms_per_step'])"
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-variants --no-parity 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('f64', d['value'], d['ms_per_step'])"
done
SWEEP_SIZES=6000,10000,25000 timeout 600 python scratch/small_template_sweep.py
} > $O 2>&1
tail -20 $O
