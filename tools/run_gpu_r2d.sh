#!/bin/bash
mkdir -p gpurun_out
{
echo "== CLI --gpus 2"
./bin/fgfa --gpus 2 -I tests/golden/ref_ex2.gfa depth -d > gpurun_out/cli_out.txt 2> gpurun_out/cli_err.txt; echo "rc=$?"
head -c 600 gpurun_out/cli_out.txt; echo ---; head -c 1500 gpurun_out/cli_err.txt
cat /etc/nccl.conf 2>/dev/null
echo "== pytest multi (2 GPUs)"
timeout 900 python -m pytest tests/test_multi_gpu_abi.py tests/test_multi_rank_gpu.py -q -m gpu 2>&1 | tail -30
} > gpurun_out/r2d2.log 2>&1
cat gpurun_out/r2d2.log
