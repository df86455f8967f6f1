#!/bin/bash
# round 2, GPU call 11: evidence -- final bench (both arms), launch list of the timed region, ncu captures, sanitizers
mkdir -p gpurun_out
O=gpurun_out/r02_run11.txt
{
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "== bench"; timeout 1500 python bench.py > gpurun_out/r02_bench11.json 2> gpurun_out/r02_bench11.err; echo "rc=$?"; tail -3 gpurun_out/r02_bench11.err
echo "== bench ref"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench11_ref.json 2>/dev/null; echo "rc=$?"
echo "== bench f32"; timeout 900 python bench.py --dtype f32 --no-cpu-baseline --no-variants > gpurun_out/r02_bench11_f32.json 2>/dev/null; echo "rc=$?"
} > $O 2>&1
# launch list of the timed region (NVTX range 'timed'), full-size workload
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" --csv \
    --log-file gpurun_out/r02_launches_raw.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-variants --no-parity > gpurun_out/r02_ncu_launches.log 2>&1
echo "launches rc=$?" >> $O
timeout 900 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "timed/" -k regex:reweight_hist -c 1 \
    -o gpurun_out/prof_r02_pair2 -f python bench.py --dtype f32 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-variants --no-parity > gpurun_out/ncu_full_r02_pair2.log 2>&1
echo "ncu pair rc=$?" >> $O
timeout 900 ncu --set full --clock-control none --profile-from-start off -o gpurun_out/prof_r02_hist -f python scratch/ncu_hist_once.py > gpurun_out/ncu_hist.log 2>&1
echo "ncu hist rc=$?" >> $O
# sanitizers over the tests of the round-2 kernels
SEL="planned or large_binning or 3200 or pair_kernel or one_call or hist_options or fp32_mode_vs or resample_to_irregular"
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 77 --print-limit 20 \
    python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/r02_sanitize_${tool}.log 2>&1
  echo "$tool rc=$?" >> $O; grep -E "passed|failed|SUMMARY" gpurun_out/r02_sanitize_${tool}.log | tail -3 >> $O
done
tail -30 $O
