#!/bin/bash
# round 2, GPU call 15: carry-save accumulators + uniform-warp path, rank quantisation, epilogue loads; small-kernel profile
mkdir -p gpurun_out
O=gpurun_out/r02_run15.txt
{
echo "== pytest (all)"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "== bench hist"; timeout 900 python scratch/bench_hist.py 2>&1 | grep 3200
echo "== large bins probe"; timeout 600 python scratch/large_bins_probe.py 1e8 f64; timeout 600 python scratch/large_bins_probe.py 1e8 f32
echo "== small template sweep"; timeout 900 python scratch/small_template_sweep.py
SWEEP_REPS=3 SWEEP_SIZES=10000 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_small_launches.csv python scratch/small_template_sweep.py > /dev/null 2>&1
python - <<'P'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r02_small_launches.csv")) if len(r) > 10]
h = rows[0]; ik, iv, ig, ib = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
seen = {}
for r in rows[1:]:
    if "reweight_hist" in r[ik] or "reduce_chi2" in r[ik]:
        seen.setdefault(r[ik][:60] + r[ig], []).append(float(r[iv]))
for k, v in seen.items(): print("%-90s median %.0f ns over %d launches" % (k, sorted(v)[len(v) // 2], len(v)))
P
} > $O 2>&1
SWEEP_REPS=1 SWEEP_SIZES=10000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:reweight_hist_kernel -s 6 -c 1 -o gpurun_out/prof_r02_small -f python scratch/small_template_sweep.py > gpurun_out/ncu_small.log 2>&1
echo "ncu small rc=$?" >> $O
tail -60 $O
