#!/bin/bash
# Window-kernel prototype on the GPU box: small configs under compute-sanitizer, then timings + oracle parity.
mkdir -p gpurun_out
out=gpurun_out/ubench_win_${1:-r2_01}.log
cfgs=${2:-"S T B C E R U"}
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
for c in S T; do
  echo "== sanitizer memcheck $c"
  UBENCH_ONLY="W r8 s3 f8" timeout 300 compute-sanitizer --tool memcheck ./build/ubench_win $c 1 1 2>&1 | grep -E "ERROR SUMMARY|Invalid|PARITY|FAIL|mismatch" | head -20
done
echo "== racecheck S"
UBENCH_ONLY="W r8 s3 f8" timeout 300 compute-sanitizer --tool racecheck ./build/ubench_win S 1 1 2>&1 | grep -E "RACECHECK SUMMARY|hazard|PARITY|FAIL" | head -10
echo "== synccheck T"
UBENCH_ONLY="W r8 s3 f8" timeout 300 compute-sanitizer --tool synccheck ./build/ubench_win T 1 1 2>&1 | grep -E "ERROR SUMMARY|Barrier|PARITY|FAIL" | head -10
for c in $cfgs; do
  echo "== $c"
  timeout 240 ./build/ubench_win $c 10 1 2>&1
done
} > $out 2>&1
tail -150 $out
