#!/bin/bash
# Standard single-GPU measurement cycle (run under gpurun): tests, bench, launch list, full capture.
# usage: scratch/gpu_cycle.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_$TAG.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['launch_ms'], d['clocks'], d['cpu_baseline']['value'], d['gpu_launches'])"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo "ref rc=$?"
# launch list of the timed region (NVTX range 'timed'), small workload
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --events-per-gpu 1.2e7 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
echo "launches rc=$?"
# one full capture of the dominant kernel
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "timed/" -k regex:reweight_hist -c 1 \
    -o gpurun_out/prof_$TAG -f python bench.py --events-per-gpu 1.2e7 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full rc=$?"
