#!/bin/bash
mkdir -p gpurun_out
{
echo "== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --no-extra 2>&1 | tail -5 | tee gpurun_out/r2_bench_n2_b.json | cut -c1-3000
echo "== CLI --gpus 2"
./bin/fgfa --gpus 2 -I tests/golden/ref_ex2.gfa depth -d | head -3
timeout 600 python -m pytest tests/test_multi_gpu_abi.py -q -m gpu -k cli 2>&1 | tail -3
} > gpurun_out/r2e.log 2>&1
cat gpurun_out/r2e.log
