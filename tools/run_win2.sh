#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/ubench_win_${1:-r2_02}.log
{
echo "== smem ops"
timeout 120 ./build/ubench_smem
for c in C R; do
  for v in "span=inf" DBG1 DBG2 DBG3; do
    echo "== $c $v"
    UBENCH_ONLY="$v" timeout 240 ./build/ubench_win $c 5 0 2>&1 | grep -v "^cfg"
  done
done
echo "== ncu"
UBENCH_ONLY="span=inf" timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_window_count -s 2 -c 1 -o gpurun_out/r2_prof_window_01 -f ./build/ubench_win C 1 0 2>&1 | tail -5
} > $out 2>&1
tail -70 $out
