#!/bin/bash
# round 2, GPU call 12: flux systematics inside the template kernel, astro_weights
mkdir -p gpurun_out
O=gpurun_out/r02_run12.txt
{
echo "== pytest (new)"; timeout 900 python -m pytest tests -m gpu -q -x -k "flux or astro or callers or one_call or engine" 2>&1 | tail -8
echo "== bench flux fold"; timeout 600 python scratch/bench_flux_fold.py 1e8 f64; timeout 600 python scratch/bench_flux_fold.py 1e8 f32
echo "== pytest (all)"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5
} > $O 2>&1
tail -40 $O
