#!/bin/bash
# round 2, GPU call 25 (8 GPUs): the exchange and the bench line at N = 8 with the final kernels
mkdir -p gpurun_out
O=gpurun_out/r02_run25.txt
{
echo "== exchange N=8"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 scratch/check_exchange.py 2>&1 | grep -v "Warning\|warn\|OMP_NUM\|\*\*\*" | tail -6
echo "== bench N=8"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench25_n8.json 2> gpurun_out/r02_bench25_n8.err; echo "rc=$?"; tail -3 gpurun_out/r02_bench25_n8.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/r02_bench25_n8.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", (d.get("parity_check") or {}).get("ok"))
for k, v in (d.get("variants") or {}).items():
    print(k, json.dumps({kk: v[kk] for kk in v if kk in ("value", "ms_per_step", "sizes")})[:600])
P
} > $O 2>&1
tail -40 $O
