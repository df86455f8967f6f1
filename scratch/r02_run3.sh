#!/bin/bash
# round 2, GPU call 3: tests, FP32-mode kernel after the shared-memory state change, the new bench line
mkdir -p gpurun_out
O=gpurun_out/r02_run3.txt
{
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== kbench_mp"
timeout 900 python scratch/kbench_mp.py '{"mp_b2": [], "mp_b3": ["-DPISAB_MP_MIN_BLOCKS=3"]}' 2>&1 | grep -v Warning
echo "== bench"; timeout 1200 python bench.py > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err; echo "rc=$?"; tail -5 gpurun_out/r02_bench3.err
echo "== bench ref"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-400
} > $O 2>&1
tail -30 $O
