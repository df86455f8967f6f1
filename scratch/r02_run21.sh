#!/bin/bash
# round 2, GPU call 21: re-check after the shared-memory attribute fix of the single-wave pair kernel
mkdir -p gpurun_out
O=gpurun_out/r02_run21.txt
{
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "== bench f32"; timeout 900 python bench.py --dtype f32 --no-cpu-baseline --no-variants > gpurun_out/r02_bench21_f32.json 2>/dev/null; echo "rc=$?"
echo "== small"; SWEEP_SIZES=3000,6000,10000,25000 timeout 600 python scratch/small_template_sweep.py
} > $O 2>&1
SEL="planned or large_binning or 3200 or pair_kernel or one_call or hist_options or flux or astro or sort_order or hypersurface or engine"
for tool in memcheck racecheck; do
  timeout 1800 compute-sanitizer --tool $tool --error-exitcode 77 --print-limit 20 \
    python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/r02_sanitize_${tool}.log 2>&1
  echo "$tool rc=$?" >> $O; grep -E "passed|failed|SUMMARY" gpurun_out/r02_sanitize_${tool}.log | tail -3 >> $O
done
tail -40 $O
