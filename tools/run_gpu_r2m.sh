#!/bin/bash
# round 2, call m (N GPUs): push form of the exchange against pull and NCCL
N=${1:-2}
mkdir -p gpurun_out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
{
nvidia-smi -L | head -8
echo "== single-device exchange tests"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "exchange" 2>&1 | tail -3
echo "== check_fused C N=$N"
run 29553 tools/check_fused_exchange.py C 2>&1 | grep -E "^\{|Error|error" | tail -3 | tee gpurun_out/r2_check_push_n$N.json
echo "== bench N=$N"
run 29551 bench.py --gpus $N 2>&1 | grep -E "^\{|Error|error" | tail -2 | tee gpurun_out/r2_bench_n${N}_push.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['ms_per_step'], d['engine'], d['exchange']); print(d['split']); print(d['parity']); print('e2e', d['e2e']['ms_per_step'], d['extra_configs'])"
} > gpurun_out/r2m_n$N.log 2>&1
cat gpurun_out/r2m_n$N.log
