#!/bin/bash
# round 2, GPU call 24: randomised-parameter parity test
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_prob3.py -m gpu -q -k "random_parameter" 2>&1 | tail -30 > gpurun_out/r02_run24.txt
tail -30 gpurun_out/r02_run24.txt
