#!/bin/bash
# ncu --set full over one launch of each HBM-bound kernel (profile range = the second pass of scratch/ncu_hbm_once.py)
mkdir -p gpurun_out
ncu --set full --clock-control none --profile-from-start off -o gpurun_out/prof_hbm -f python scratch/ncu_hbm_once.py > gpurun_out/ncu_hbm.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_hbm.log
ncu -i gpurun_out/prof_hbm.ncu-rep --page raw --csv > gpurun_out/prof_hbm_raw.csv 2>/dev/null; wc -l gpurun_out/prof_hbm_raw.csv
