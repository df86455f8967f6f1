#!/bin/bash
# round 2, GPU call 16: interleaved chunk mapping for single-wave grids
mkdir -p gpurun_out
O=gpurun_out/r02_run16.txt
{
echo "== pytest (all)"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "== small template sweep"; timeout 900 python scratch/small_template_sweep.py
SWEEP_REPS=3 SWEEP_SIZES=10000 timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_active.min,sm__cycles_active.avg,sm__cycles_active.max --clock-control none --csv --log-file gpurun_out/r02_small_launches.csv python scratch/small_template_sweep.py > /dev/null 2>&1
python - <<'P'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r02_small_launches.csv")) if len(r) > 10]
h = rows[0]; ik, im, iv, ig = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Grid Size")
seen = {}
for r in rows[1:]:
    if "reweight_hist" in r[ik] or "reduce_chi2" in r[ik]:
        seen.setdefault((r[ik][:50] + r[ig], r[im]), []).append(float(r[iv].replace(",", "")))
for k, v in seen.items(): print("%-70s %-28s median %.0f over %d launches" % (k[0], k[1], sorted(v)[len(v) // 2], len(v)))
P
echo "== bench_small"; timeout 600 python scratch/bench_small.py 2>&1 | tail -8
} > $O 2>&1
tail -60 $O
