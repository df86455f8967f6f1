#!/bin/bash
# round 2, GPU call 10 (8 GPUs): the exchange and the bench line at N = 8
mkdir -p gpurun_out
O=gpurun_out/r02_run10.txt
{
echo "== exchange N=8"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 scratch/check_exchange.py 2>&1 | grep -v "Warning\|warn\|OMP_NUM\|\*\*\*" | tail -16
echo "== bench N=8"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench10_n8.json 2> gpurun_out/r02_bench10_n8.err; echo "rc=$?"; tail -3 gpurun_out/r02_bench10_n8.err
} > $O 2>&1
tail -40 $O
