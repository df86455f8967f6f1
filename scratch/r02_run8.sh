#!/bin/bash
# round 2, GPU call 8: pair-kernel variants
mkdir -p gpurun_out
O=gpurun_out/r02_run8.txt
{
echo "== kbench_pair"
timeout 1500 python scratch/kbench_pair.py '{"lockstep": [], "no_lockstep": ["-DPISAB_PAIR_NO_LOCKSTEP"], "lockstep_b3": ["-DPISAB_PAIR_MIN_BLOCKS=3"], "no_lockstep_b3": ["-DPISAB_PAIR_NO_LOCKSTEP", "-DPISAB_PAIR_MIN_BLOCKS=3"], "lockstep_128x4": ["-DPISAB_BLOCK=128", "-DPISAB_PAIR_MIN_BLOCKS=4", "-DPISAB_MIN_BLOCKS=4", "-DPISAB_MP_MIN_BLOCKS=6"]}' 2>&1 | grep -v Warning
} > $O 2>&1
tail -40 $O
