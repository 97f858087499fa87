#!/bin/bash
# round 2, call k: GPU test suite + bench with the OVL kernel W; ncu full capture of it; launch list of one bench step
mkdir -p gpurun_out
{
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "== bench N=1"
timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/r2_bench_n1_ovl.json | cut -c1-1500
echo "== ncu full, kernel W (OVL)"
UBENCH_ONLY="OVL W r8" timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_window_count -s 2 -c 1 -o gpurun_out/r2_prof_window_04_ovl -f ./build/ubench_win C 1 0 2>&1 | tail -3
echo "== ncu launch list of the bench"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_ovl.csv python bench.py --steps 2 --warmup 1 --no-extra > gpurun_out/r2_launches_bench.log 2>&1
tail -2 gpurun_out/r2_launches_bench.log | cut -c1-300
} > gpurun_out/r2k.log 2>&1
cat gpurun_out/r2k.log
