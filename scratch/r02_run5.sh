#!/bin/bash
# round 2, GPU call 5 (2 GPUs): all GPU tests on one GPU, then the peer-memory exchange and the bench at N = 2
mkdir -p gpurun_out
O=gpurun_out/r02_run5.txt
{
echo "== pytest"; CUDA_VISIBLE_DEVICES=0 timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25
echo "== exchange"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scratch/check_exchange.py 2>&1 | grep -v "Warning\|warn" | tail -20
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench5_n2.json 2> gpurun_out/r02_bench5_n2.err; echo "rc=$?"; tail -3 gpurun_out/r02_bench5_n2.err
} > $O 2>&1
tail -60 $O
