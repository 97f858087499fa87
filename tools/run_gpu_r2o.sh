#!/bin/bash
# round 2, call o: ncu full capture of the final kernel W + pre-pass + B2, launch list of one bench run
mkdir -p gpurun_out
{
echo "== ncu full, final kernel W"
UBENCH_ONLY="OVL W r8" timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_window_count -s 2 -c 1 -o gpurun_out/r2_prof_window_05_final -f ./build/ubench_win C 1 0 2>&1 | tail -2
echo "== ncu full, pre-pass + B2"
UBENCH_ONLY="OVL W r8" timeout 600 ncu --set full --clock-control none -k regex:"k_bin|k_uniq_from" -s 8 -c 4 -o gpurun_out/r2_prof_prepass_02_final -f ./build/ubench_win C 1 0 2>&1 | tail -2
echo "== ncu launch list of the bench"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 2 --warmup 1 --no-extra > gpurun_out/r2_launches_bench.log 2>&1
tail -1 gpurun_out/r2_launches_bench.log | cut -c1-200
} > gpurun_out/r2o.log 2>&1
cat gpurun_out/r2o.log
