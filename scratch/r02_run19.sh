#!/bin/bash
# round 2, GPU call 19: coalesced in-kernel epilogue, A/B against the two-launch form
mkdir -p gpurun_out
O=gpurun_out/r02_run19.txt
export SWEEP_SIZES=3000,6000,10000,25000,100000 SWEEP_REPS=200
{
echo "== pytest (all)"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
for cfg in "" "PISAB_NO_FUSED_EPILOGUE=1"; do
  echo "== [$cfg] sweep"; env $cfg timeout 600 python scratch/small_template_sweep.py
  echo "== [$cfg] bench f64 / f32"
  env $cfg timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-variants --no-parity 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('f64', d['value'], d['ms_per_step'])"
  env $cfg timeout 600 python bench.py --dtype f32 --no-cpu-baseline --no-e2e --no-variants --no-parity 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('f32', d['value'], d['ms_per_step'])"
done
} > $O 2>&1
tail -80 $O
