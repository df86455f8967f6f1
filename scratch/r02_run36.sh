#!/bin/bash
# round 2, GPU call 36: synccheck + initcheck over the decay / scan / layers tests
mkdir -p gpurun_out
O=gpurun_out/r02_run36.txt
SEL="decay or scan or layers or prem59"
: > $O
for tool in synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 77 --print-limit 20 \
    python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/r02_sanitize_decay_${tool}.log 2>&1
  echo "$tool rc=$?" >> $O; grep -E "passed|failed|SUMMARY" gpurun_out/r02_sanitize_decay_${tool}.log | tail -3 >> $O
done
cat $O
