#!/bin/bash
# round 2, GPU call 27: number of whole waves of the multi-wave grid
mkdir -p gpurun_out
O=gpurun_out/r02_run27.txt
{
for w in 32 16 24 48 64 32; do
echo "[waves $w]"
PISAB_EXP_WAVES=$w timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-variants --no-parity 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('f64', d['value'], d['ms_per_step'])"
PISAB_EXP_WAVES=$w timeout 600 python bench.py --dtype f32 --no-cpu-baseline --no-e2e --no-variants --no-parity 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('f32', d['value'], d['ms_per_step'])"
done
} > $O 2>&1
tail -24 $O
