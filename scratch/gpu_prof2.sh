#!/bin/bash
# two full captures of the fused kernel: standard matter and NSI (general path)
TAG=${1:-x}
for mode in std nsi; do
  extra=""; [ $mode = nsi ] && extra="--nsi"
  ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "timed/" -k regex:reweight_hist -c 1 \
      -o gpurun_out/prof_${TAG}_$mode -f python bench.py --events-per-gpu 1.2e7 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline $extra > gpurun_out/ncu_full_${TAG}_$mode.log 2>&1
  echo "$mode rc=$?"
done
