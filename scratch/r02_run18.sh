#!/bin/bash
# round 2, GPU call 18: A/B of the fused epilogue and the snake order
mkdir -p gpurun_out
O=gpurun_out/r02_run18.txt
export SWEEP_SIZES=6000,10000,25000 SWEEP_REPS=200
{
for cfg in "" "PISAB_NO_FUSED_EPILOGUE=1" "PISAB_NO_SNAKE=1" "PISAB_NO_FUSED_EPILOGUE=1 PISAB_NO_SNAKE=1" "PISAB_NO_FUSED_EPILOGUE=1 PISAB_NO_INTERLEAVE=1"; do
  echo "== [$cfg] sweep"; env $cfg timeout 600 python scratch/small_template_sweep.py
  echo "== [$cfg] bench f64 / f32"
  env $cfg timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-variants --no-parity 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('f64', d['value'], d['ms_per_step'])"
  env $cfg timeout 600 python bench.py --dtype f32 --no-cpu-baseline --no-e2e --no-variants --no-parity 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('f32', d['value'], d['ms_per_step'])"
done
} > $O 2>&1
tail -80 $O
