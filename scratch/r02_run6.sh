#!/bin/bash
# round 2, GPU call 6: all GPU tests after the histogram / exchange / epilogue work; scan latency
mkdir -p gpurun_out
O=gpurun_out/r02_run6.txt
{
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25
echo "== bench_small"; timeout 300 python scratch/bench_small.py 2>&1 | grep -v Warning | tail -12
echo "== bench_scan"; timeout 600 python scratch/bench_scan.py 2>&1 | grep -v Warning | tail -12
} > $O 2>&1
tail -60 $O
