#!/bin/bash
# round 2, GPU call 23 (2 GPUs): tests, the exchange, the bench line at N = 2 with the final kernels
mkdir -p gpurun_out
O=gpurun_out/r02_run23.txt
{
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5
echo "== exchange N=2"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scratch/check_exchange.py 2>&1 | grep -v "Warning\|warn\|OMP_NUM\|\*\*\*" | tail -12
echo "== sharded pipeline N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 scratch/check_sharded_pipeline.py 2>&1 | grep -v "Warning\|warn\|OMP_NUM\|\*\*\*" | tail -8
echo "== bench N=2"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench23_n2.json 2> gpurun_out/r02_bench23_n2.err; echo "rc=$?"; tail -3 gpurun_out/r02_bench23_n2.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/r02_bench23_n2.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", (d.get("parity_check") or {}).get("ok"))
for k, v in (d.get("variants") or {}).items():
    print(k, json.dumps({kk: v[kk] for kk in v if kk in ("value", "ms_per_step", "parity_check", "sizes")})[:500])
P
} > $O 2>&1
tail -50 $O
