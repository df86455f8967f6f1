#!/bin/bash
# round 2, GPU call 30: neutrino-decay branch -- parity tests, timing, memcheck
mkdir -p gpurun_out
O=gpurun_out/r02_run30.txt
{
echo "== pytest decay"; timeout 900 python -m pytest tests -m gpu -q -x -k "decay or golden_pickles or stage_contract" 2>&1 | tail -15
echo "== bench decay"; timeout 600 python scratch/bench_decay.py 2>&1 | tail -20
echo "== memcheck"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 77 --print-limit 20 python -m pytest tests -m gpu -x -q -k "fused_reweight_hist_with_neutrino_decay or decay_reference_fixture" > gpurun_out/r02_sanitize_decay_memcheck.log 2>&1; echo "rc=$?"; grep -E "passed|failed|SUMMARY" gpurun_out/r02_sanitize_decay_memcheck.log | tail -3
} > $O 2>&1
tail -60 $O
