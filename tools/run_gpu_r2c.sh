#!/bin/bash
mkdir -p gpurun_out
{
echo "== pytest multi + window"
timeout 900 python -m pytest tests/test_multi_gpu_abi.py tests/test_window_engine.py -x -q -m gpu 2>&1 | tail -8
echo "== bench"
timeout 900 python bench.py 2>&1 | tail -3 | tee gpurun_out/r2_bench_n1.json
} > gpurun_out/r2c.log 2>&1
cat gpurun_out/r2c.log
