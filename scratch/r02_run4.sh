#!/bin/bash
# round 2, GPU call 4: new tests (callers, planned / exact histograms, reference cfg), histogram kernels, MP kernel, bench
mkdir -p gpurun_out
O=gpurun_out/r02_run4.txt
{
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== bench_hist"; timeout 600 python scratch/bench_hist.py 2>&1 | grep -v Warning
echo "== kbench_mp"; timeout 600 python scratch/kbench_mp.py '{"mp_default": []}' 2>&1 | grep -v Warning
echo "== bench"; timeout 1200 python bench.py --no-cpu-baseline > gpurun_out/r02_bench4.json 2> gpurun_out/r02_bench4.err; echo "rc=$?"; tail -5 gpurun_out/r02_bench4.err
} > $O 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "timed/" -k regex:reweight_hist -c 1 \
    -o gpurun_out/prof_r02_mp2 -f python bench.py --dtype f32 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-variants --no-parity > gpurun_out/ncu_full_r02_mp2.log 2>&1
echo "ncu rc=$?" >> $O
tail -40 $O
