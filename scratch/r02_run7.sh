#!/bin/bash
# round 2, GPU call 7: the two-events-per-thread FP32 kernel -- tests, bench --dtype f32, ncu capture, full bench
mkdir -p gpurun_out
O=gpurun_out/r02_run7.txt
{
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15
echo "== bench f32"; timeout 900 python bench.py --dtype f32 --no-cpu-baseline --no-variants 2>/dev/null | cut -c1-1500
echo "== bench"; timeout 1200 python bench.py --no-cpu-baseline > gpurun_out/r02_bench7.json 2> gpurun_out/r02_bench7.err; echo "rc=$?"; tail -3 gpurun_out/r02_bench7.err
echo "== bench_scan"; timeout 600 python scratch/bench_scan.py 2>&1 | grep -v Warning | tail -8
} > $O 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "timed/" -k regex:reweight_hist -c 1 \
    -o gpurun_out/prof_r02_pair -f python bench.py --dtype f32 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-variants --no-parity > gpurun_out/ncu_full_r02_pair.log 2>&1
echo "ncu rc=$?" >> $O
tail -40 $O
