#!/bin/bash
# round 2, GPU call 38: pair kernel with the block size picked per grid (64 x 8 for multi-wave standard matter): tests, timing, bench f32
mkdir -p gpurun_out
O=gpurun_out/r02_run38.txt
{
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -x -k "pair or fp32 or f32 or flux or float32 or callers or fused" 2>&1 | tail -3
echo "== kbench f32"; python scratch/kbench_std.py 96000000 f32 2>&1 | tail -2
echo "== small templates"; SWEEP_SIZES=6000,10000,25000,100000 python scratch/small_template_sweep.py 2>&1 | grep float32
echo "== bench f32"; timeout 600 python bench.py --dtype f32 --no-cpu-baseline --no-variants > gpurun_out/r02_bench38_f32.json 2>/dev/null; echo rc=$?
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench38_f32.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parity_check'])"
} > $O 2>&1
cat $O
