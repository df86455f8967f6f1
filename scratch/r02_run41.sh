#!/bin/bash
# round 2, GPU call 41: ncu --set full of the FP32-mode template kernel as shipped (64 x 8 blocks for the bench's grid)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "timed/" -k regex:reweight_hist -c 1 \
    -o gpurun_out/prof_r02_final_f32b -f python bench.py --dtype f32 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-variants --no-parity > gpurun_out/ncu_full_r02_final_f32b.log 2>&1
echo "f32 rc=$?"
