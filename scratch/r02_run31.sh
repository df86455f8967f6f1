#!/bin/bash
# round 2, GPU call 31: decay in the scan, decay tests, ncu capture of the decay template kernel
mkdir -p gpurun_out
O=gpurun_out/r02_run31.txt
{
echo "== pytest decay + scan"; timeout 900 python -m pytest tests -m gpu -q -x -k "decay or scan or golden_pickles" 2>&1 | tail -6
echo "== bench decay"; timeout 600 python scratch/bench_decay.py 2>&1 | grep float64 | head -4
} > $O 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:reweight_hist_decay -s 2 -c 1 \
    -o gpurun_out/prof_r02_decay -f python scratch/ncu_decay_once.py > gpurun_out/ncu_full_r02_decay.log 2>&1
echo "ncu rc=$?" >> $O
tail -30 $O
