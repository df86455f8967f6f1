#!/bin/bash
# round 2, GPU call 39: final state -- all GPU tests, smoke, both bench arms, memcheck + racecheck over the FP32-mode / pair-kernel tests
mkdir -p gpurun_out
O=gpurun_out/r02_run39.txt
{
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench"; timeout 1500 python bench.py > gpurun_out/r02_bench39.json 2> gpurun_out/r02_bench39.err; echo "rc=$?"
echo "== bench ref"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench39_ref.json 2>/dev/null; echo "rc=$?"
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 77 --print-limit 20 \
    python -m pytest tests -m gpu -x -q -k "pair_kernel or fp32_mode or flux_systematics_folded" > gpurun_out/r02_sanitize_pair_${tool}.log 2>&1
  echo "$tool rc=$?"; grep -E "passed|failed|SUMMARY" gpurun_out/r02_sanitize_pair_${tool}.log | tail -3
done
} > $O 2>&1
cat $O
