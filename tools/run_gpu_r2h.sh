#!/bin/bash
# round 2, call h: BITROWS form of kernel W (seen-bits as global bitmap rows via run heads) against the mask form
mkdir -p gpurun_out
{
for c in S T; do
  echo "== $c"
  UBENCH_ONLY="BITROWS" timeout 120 ./build/ubench_win $c 3 1 2>&1 | grep -E "total best|PARITY|FAIL|err"
done
for c in C E R U; do
  echo "== $c"
  UBENCH_ONLY="BITROWS" timeout 300 ./build/ubench_win $c 8 1 2>&1 | grep -E "total best|PARITY|FAIL|err|S1"
  UBENCH_ONLY="W r8 s2" timeout 300 ./build/ubench_win $c 8 0 2>&1 | grep -E "total best"
done
} > gpurun_out/r2h.log 2>&1
cat gpurun_out/r2h.log
