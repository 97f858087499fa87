#!/bin/bash
# round 2, call l: programmatic dependent launch on/off
mkdir -p gpurun_out
{
echo "== bench N=1 PDL on"
timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/r2_bench_n1_pdl.json
python -c "import json,sys; d=json.loads(open('gpurun_out/r2_bench_n1_pdl.json').read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['whole_step_frac'], d['split']['depth_only_ms_per_step'], d['parity'], {k:(v['ms_per_step'],v['parity_vs_oracle']) for k,v in d['extra_configs'].items()})" || head -c 600 gpurun_out/r2_bench_n1_pdl.json
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
} > gpurun_out/r2l.log 2>&1
cat gpurun_out/r2l.log
