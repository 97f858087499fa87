#!/bin/bash
mkdir -p gpurun_out
{
for g in 32 64 128; do
echo "== ubench C L2 fetch $g"
UBENCH_L2_FETCH=$g UBENCH_ONLY="W r8 s2" timeout 240 ./build/ubench_win C 10 1 2>&1 | grep -E "total best|mismatch|S1|Granul"
done
} > gpurun_out/r2b.log 2>&1
cat gpurun_out/r2b.log
