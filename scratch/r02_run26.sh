#!/bin/bash
# round 2, GPU call 26: whole waves for the multi-wave grid (789 instead of 790 ranks per container), A/B on one box
mkdir -p gpurun_out
O=gpurun_out/r02_run26.txt
{
for cfg in "" "PISAB_EXP_CEIL=1" "" "PISAB_EXP_CEIL=1"; do
echo "[$cfg]"
env $cfg timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-variants --no-parity 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('f64', d['value'], d['ms_per_step'])"
env $cfg timeout 600 python bench.py --dtype f32 --no-cpu-baseline --no-e2e --no-variants --no-parity 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('f32', d['value'], d['ms_per_step'])"
done
} > $O 2>&1
tail -20 $O
