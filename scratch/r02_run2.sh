#!/bin/bash
# round 2, GPU call 2: FP32-mode (mixed precision) kernel -- tests, variants, accuracy, bench lines
mkdir -p gpurun_out
O=gpurun_out/r02_run2.txt
{
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== kbench_mp"
timeout 1500 python scratch/kbench_mp.py '{"mp_b2": [], "mp_b3": ["-DPISAB_MP_MIN_BLOCKS=3"], "mp_b4": ["-DPISAB_MP_MIN_BLOCKS=4"], "mp_b128x4": ["-DPISAB_BLOCK=128", "-DPISAB_MP_MIN_BLOCKS=4", "-DPISAB_MIN_BLOCKS=4"]}' 2>&1 | grep -v Warning
echo "== bench f32 mixed"; timeout 600 python bench.py --dtype f32 --no-cpu-baseline 2>/dev/null
echo "== bench f32 fp64math"; timeout 600 python bench.py --dtype f32 --f32-math fp64 --no-e2e --no-cpu-baseline 2>/dev/null
} > $O 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "timed/" -k regex:reweight_hist -c 1 \
    -o gpurun_out/prof_r02_mp -f python bench.py --dtype f32 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_r02_mp.log 2>&1
echo "ncu rc=$?" >> $O
tail -30 $O
