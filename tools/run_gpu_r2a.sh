#!/bin/bash
mkdir -p gpurun_out
{
echo "== pytest window"
timeout 900 python -m pytest tests/test_window_engine.py -x -q -m gpu 2>&1 | tail -15
echo "== pytest parity"
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -8
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -3
echo "== ncu prepass"
UBENCH_ONLY="W r8 s2" timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_bin -s 6 -c 3 -o gpurun_out/r2_prof_prepass_01 -f ./build/ubench_win C 1 0 2>&1 | tail -3
} > gpurun_out/r2a.log 2>&1
cat gpurun_out/r2a.log
