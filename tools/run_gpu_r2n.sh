#!/bin/bash
# round 2, call n (N GPUs): timing of the three exchanges on config C (check tool) + bench without extras
N=${1:-8}
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
{
echo "== check_fused C N=$N"
run 29553 tools/check_fused_exchange.py C 2>&1 | grep -E "^\{|Error|error" | tail -3 | tee gpurun_out/r2_check_push_n$N.json
echo "== bench N=$N --no-extra"
run 29551 bench.py --gpus $N --no-extra 2>&1 | grep -E "^\{|Error|error" | tail -2 | tee gpurun_out/r2_bench_n${N}_push.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['ms_per_step'], d['engine'], d['exchange']); print(d['split']); print(d['parity']); print('e2e', d['e2e']['ms_per_step'])"
} > gpurun_out/r2n_n$N.log 2>&1
cat gpurun_out/r2n_n$N.log
