#!/bin/bash
mkdir -p gpurun_out
{
for c in C R; do
for pm in 0 1; do
for v in "W r8 s2" "DESC-AHEAD W r8 s2" "DESC-AHEAD W r8 s3"; do
if [ $pm = 1 ]; then export UBENCH_PERMUTE=1; else unset UBENCH_PERMUTE; fi
echo -n "permute=$pm $c: "
UBENCH_ONLY="$v" timeout 240 ./build/ubench_win $c 8 1 2>&1 | grep -E "total best|FAIL"
done
done
done
} > gpurun_out/r2f.log 2>&1
cat gpurun_out/r2f.log
