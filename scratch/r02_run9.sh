#!/bin/bash
# round 2, GPU call 9: pair-kernel block shapes; tests
mkdir -p gpurun_out
O=gpurun_out/r02_run9.txt
{
echo "== kbench_pair"
timeout 1500 python scratch/kbench_pair.py '{"b128x4": [], "b128x5": ["-DPISAB_PAIR_MIN_BLOCKS=5"], "b64x8": ["-DPISAB_PAIR_BLOCK=64", "-DPISAB_PAIR_MIN_BLOCKS=8"], "b256x2": ["-DPISAB_PAIR_BLOCK=256", "-DPISAB_PAIR_MIN_BLOCKS=2"], "b192x2": ["-DPISAB_PAIR_BLOCK=192", "-DPISAB_PAIR_MIN_BLOCKS=2"], "b96x5": ["-DPISAB_PAIR_BLOCK=96", "-DPISAB_PAIR_MIN_BLOCKS=5"]}' 2>&1 | grep -v Warning
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5
echo "== bench f32"; timeout 900 python bench.py --dtype f32 --no-cpu-baseline --no-variants --no-e2e 2>/dev/null | cut -c1-300
} > $O 2>&1
tail -40 $O
