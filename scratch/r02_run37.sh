#!/bin/bash
# round 2, GPU call 37: block size of the FP32-mode pair kernel (128 x 4, 64 x 8, 32 x 16) at 9.6e7 events and on analysis-size templates
mkdir -p gpurun_out
O=gpurun_out/r02_run37.txt
{
python scratch/kbench_std.py 96000000 f32 2>&1 | tail -8
for lib in "" scratch/variants/libpisa_b64x8.so scratch/variants/libpisa_b32x16.so; do
  echo "== small templates, lib=[$lib]"
  PISAB_LIB=$lib SWEEP_SIZES=3000,6000,10000,25000,100000 python scratch/small_template_sweep.py 2>&1 | grep float32
done
} > $O 2>&1
cat $O
