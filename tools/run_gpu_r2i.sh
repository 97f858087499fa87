#!/bin/bash
# round 2, call i: OVL form of kernel W (loads issued before the ATOMS burst; descriptors by batch + shuffle)
mkdir -p gpurun_out
{
for c in S T; do
  echo "== $c"
  UBENCH_ONLY="OVL" timeout 120 ./build/ubench_win $c 3 1 2>&1 | grep -E "total best|PARITY|FAIL|err"
done
for c in C E R; do
  echo "== $c"
  UBENCH_ONLY="OVL" timeout 300 ./build/ubench_win $c 8 1 2>&1 | grep -E "total best|PARITY|FAIL|err"
  UBENCH_ONLY="W r8 s2" timeout 300 ./build/ubench_win $c 8 0 2>&1 | grep -E "total best" | head -1
  UBENCH_ONLY="depth-only r8 s3" timeout 300 ./build/ubench_win $c 8 0 2>&1 | grep -E "total best" | head -1
done
} > gpurun_out/r2i.log 2>&1
cat gpurun_out/r2i.log
