#!/bin/bash
# round 2, call g: load-path microbenchmark (tools/ubench_ld.cu) + the DBG variants of kernel W
mkdir -p gpurun_out
{
for c in C R; do
  echo "== ubench_ld $c"
  timeout 300 ./build/ubench_ld $c 8 2>&1
done
echo "== ubench_win C DBG variants"
UBENCH_ONLY="DBG" timeout 300 ./build/ubench_win C 8 0 2>&1 | grep -E "total best"
UBENCH_ONLY="W r8 s2" timeout 300 ./build/ubench_win C 8 0 2>&1 | grep -E "total best"
} > gpurun_out/r2g.log 2>&1
cat gpurun_out/r2g.log
