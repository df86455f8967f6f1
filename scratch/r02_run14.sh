#!/bin/bash
# round 2, GPU call 14: large-binning launch breakdown, chunked sorted kernel, small-template sweep
mkdir -p gpurun_out
O=gpurun_out/r02_run14.txt
{
echo "== pytest hist"; timeout 900 python -m pytest tests/test_gpu_hist.py -m gpu -q 2>&1 | tail -5
echo "== bench hist"; timeout 900 python scratch/bench_hist.py 2>&1 | grep 3200
echo "== large bins probe"; timeout 600 python scratch/large_bins_probe.py 1e8 f64
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_large_launches.csv python scratch/large_bins_probe.py 1e8 f64 > /dev/null 2>&1
python - <<'P'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02_large_launches.csv")) if len(r) > 10]
h = rows[0]; ik, iv = h.index("Kernel Name"), h.index("Metric Value")
last = rows[-40:]
for r in last[-12:]:
    print("%-90s %s" % (r[ik][:90], r[iv]))
P
echo "== small template sweep"; timeout 900 python scratch/small_template_sweep.py
SWEEP_REPS=3 SWEEP_SIZES=10000 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_small_launches.csv python scratch/small_template_sweep.py > /dev/null 2>&1
python - <<'P'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r02_small_launches.csv")) if len(r) > 10]
h = rows[0]; ik, iv, ig, ib = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
for r in rows[1:]:
    if "reweight_hist" in r[ik] or "reduce_chi2" in r[ik]:
        print("%-80s grid %s block %s  %s ns" % (r[ik][:80], r[ig], r[ib], r[iv]))
P
} > $O 2>&1
tail -80 $O
