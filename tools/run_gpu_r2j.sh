#!/bin/bash
# round 2, call j: ring form of kernel W (per-warp TMA ring)
mkdir -p gpurun_out
{
for c in S T; do
  echo "== $c"
  UBENCH_ONLY="RING" timeout 120 ./build/ubench_win $c 3 1 2>&1 | grep -E "total best|PARITY|FAIL|err"
done
echo "== memcheck T"
UBENCH_ONLY="RING D=2 W" timeout 300 compute-sanitizer --tool memcheck ./build/ubench_win T 1 1 2>&1 | grep -E "PARITY|FAIL|ERROR SUMMARY|Invalid|error" | head
echo "== racecheck T"
UBENCH_ONLY="RING D=1 W" timeout 300 compute-sanitizer --tool racecheck ./build/ubench_win T 1 1 2>&1 | grep -E "PARITY|FAIL|RACECHECK SUMMARY|hazard" | head
for c in C E R; do
  echo "== $c"
  UBENCH_ONLY="RING" timeout 300 ./build/ubench_win $c 8 1 2>&1 | grep -E "total best|PARITY|FAIL|err"
  UBENCH_ONLY="OVL W r8" timeout 300 ./build/ubench_win $c 8 0 2>&1 | grep -E "total best" | head -1
done
} > gpurun_out/r2j.log 2>&1
cat gpurun_out/r2j.log
