#!/bin/bash
# round 2, GPU call 20: evidence after the flux fold / large-binning / small-template work
mkdir -p gpurun_out
O=gpurun_out/r02_run20.txt
{
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "== bench"; timeout 1500 python bench.py > gpurun_out/r02_bench20.json 2> gpurun_out/r02_bench20.err; echo "rc=$?"; tail -3 gpurun_out/r02_bench20.err
echo "== bench ref"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench20_ref.json 2>/dev/null; echo "rc=$?"
echo "== bench f32"; timeout 900 python bench.py --dtype f32 --no-cpu-baseline --no-variants > gpurun_out/r02_bench20_f32.json 2>/dev/null; echo "rc=$?"
echo "== hist kernels"; timeout 900 python scratch/bench_hist.py
echo "== flux fold"; timeout 600 python scratch/bench_flux_fold.py 1e8 f64; timeout 600 python scratch/bench_flux_fold.py 1e8 f32
echo "== small templates"; timeout 600 python scratch/bench_small.py 2>&1 | tail -6
} > $O 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" --csv \
    --log-file gpurun_out/r02_launches_raw.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-variants --no-parity > gpurun_out/r02_ncu_launches.log 2>&1
echo "launches rc=$?" >> $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:reweight_hist_kernel -s 3 -c 1 \
    -o gpurun_out/prof_r02_large -f python scratch/large_bins_probe.py 1e8 f64 > gpurun_out/ncu_full_r02_large.log 2>&1
echo "ncu large rc=$?" >> $O
SEL="planned or large_binning or 3200 or pair_kernel or one_call or hist_options or flux or astro or sort_order or hypersurface or engine"
for tool in memcheck racecheck; do
  timeout 1800 compute-sanitizer --tool $tool --error-exitcode 77 --print-limit 20 \
    python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/r02_sanitize_${tool}.log 2>&1
  echo "$tool rc=$?" >> $O; grep -E "passed|failed|SUMMARY" gpurun_out/r02_sanitize_${tool}.log | tail -3 >> $O
done
tail -60 $O
