#!/bin/bash
# compute-sanitizer passes over the GPU tests (run under gpurun).  Output: gpurun_out/sanitize_*.log
mkdir -p gpurun_out
SEL="not full_size and not large and not properties"
for tool in memcheck racecheck synccheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 77 --print-limit 20 \
    python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/sanitize_${tool}.log 2>&1
  echo "$tool rc=$?"; grep -E "passed|failed|SUMMARY" gpurun_out/sanitize_${tool}.log | tail -3
done
