#!/bin/bash
# round 2, GPU call 1: issue-rate micro-benchmark, the numba reference on the box's cores, sanity tests, 1e8-event ncu capture
mkdir -p gpurun_out
O=gpurun_out/r02_run1.txt
{
echo "== host"; nproc; lscpu | grep -E "Model name|Socket|Thread|Core" ; free -g | head -2
echo "== pipe_mix"; timeout 120 scratch/micro/pipe_mix
echo "== numba chain"
for args in "--target parallel --ftype fp64 --events 1200000" "--target parallel --ftype fp32 --events 1200000" "--target cpu --ftype fp64 --events 60000" "--target parallel --ftype fp64 --events 1200000 --nsi"; do
  timeout 600 python baseline/numba_chain.py $args 2>/dev/null
done
echo "== pytest"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== bench"; timeout 900 python bench.py --no-cpu-baseline 2>/dev/null
} > $O 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "timed/" -k regex:reweight_hist -c 1 \
    -o gpurun_out/prof_r02_head -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_r02_head.log 2>&1
echo "ncu rc=$?" >> $O
tail -40 $O
