#!/bin/bash
# final evidence of the round at N GPUs: (N = 1) the GPU test suite and bench.py both arms; (N > 1) bench.py
N=${1:-1}
mkdir -p gpurun_out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
{
if [ "$N" = 1 ]; then
  echo "== pytest -m gpu"
  timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
  echo "== smoke"
  timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
  echo "== bench reference arm"
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/r2_final_reference.json | cut -c1-400
  echo "== bench N=1"
  timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/r2_final_n1.json
else
  echo "== bench N=$N"
  run 29551 bench.py --gpus $N 2>&1 | grep -E "^\{" | tail -1 > gpurun_out/r2_final_n$N.json
fi
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_final_n$N.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['engine'], d['exchange'])
print('roofline', {k:d['roofline'][k] for k in ('frac','whole_step_frac','kernel_ms','kernel')})
print('split', {k:v for k,v in d['split'].items() if k!='per_step_rank0'})
print('parity', d['parity']['vs_oracle'], 'e2e', d['e2e']['ms_per_step'], d['e2e'].get('pageable',{}).get('ms_per_step'))
print('extra', {k:(v['ms_per_step'], v['whole_step_frac'], v['parity_vs_oracle'], v['engine']) for k,v in d['extra_configs'].items()})
print('cpu', d.get('cpu_baseline') and d['cpu_baseline']['value'], 'clocks', d['clocks'])
PY
} > gpurun_out/r2_final_n$N.log 2>&1
cat gpurun_out/r2_final_n$N.log
