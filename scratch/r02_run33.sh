#!/bin/bash
# round 2, GPU call 33: final evidence (with the decay branch) -- tests, both bench arms, f32, launch list, sanitizers
mkdir -p gpurun_out
O=gpurun_out/r02_run33.txt
{
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 1500 python bench.py > gpurun_out/r02_bench33.json 2> gpurun_out/r02_bench33.err; echo "rc=$?"; tail -3 gpurun_out/r02_bench33.err
echo "== bench ref"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench33_ref.json 2>/dev/null; echo "rc=$?"
echo "== bench f32"; timeout 900 python bench.py --dtype f32 --no-cpu-baseline --no-variants > gpurun_out/r02_bench33_f32.json 2>/dev/null; echo "rc=$?"
} > $O 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" --csv \
    --log-file gpurun_out/r02_launches_raw.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-variants --no-parity > gpurun_out/r02_ncu_launches.log 2>&1
echo "launches rc=$?" >> $O
SEL="planned or large_binning or 3200 or pair_kernel or one_call or hist_options or flux or astro or sort_order or hypersurface or engine or histogram or decay"
for tool in memcheck racecheck; do
  timeout 1800 compute-sanitizer --tool $tool --error-exitcode 77 --print-limit 20 \
    python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/r02_sanitize_${tool}.log 2>&1
  echo "$tool rc=$?" >> $O; grep -E "passed|failed|SUMMARY" gpurun_out/r02_sanitize_${tool}.log | tail -3 >> $O
done
tail -40 $O
