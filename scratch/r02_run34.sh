#!/bin/bash
# round 2, GPU call 34: ncu --set full captures of the template kernels of the FINAL build (FP64 and FP32 mode), 1e8-event launches
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "timed/" -k regex:reweight_hist -c 1 \
    -o gpurun_out/prof_r02_final_f64 -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-variants --no-parity > gpurun_out/ncu_full_r02_final_f64.log 2>&1
echo "f64 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "timed/" -k regex:reweight_hist -c 1 \
    -o gpurun_out/prof_r02_final_f32 -f python bench.py --dtype f32 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-variants --no-parity > gpurun_out/ncu_full_r02_final_f32.log 2>&1
echo "f32 rc=$?"
