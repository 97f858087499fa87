#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/ubench_win_${1:-r2_06}.log
{
for c in S T; do
  echo "== sanitizer memcheck $c"
  UBENCH_ONLY="W r8 s3" timeout 300 compute-sanitizer --tool memcheck ./build/ubench_win $c 1 1 2>&1 | grep -E "ERROR SUMMARY|Invalid|PARITY|FAIL" | head -20
done
echo "== racecheck S"
UBENCH_ONLY="W r8 s3" timeout 300 compute-sanitizer --tool racecheck ./build/ubench_win S 1 1 2>&1 | grep -E "RACECHECK SUMMARY|hazard|PARITY|FAIL" | head -10
for c in S T R C E B U; do
    echo "== $c"
    timeout 240 ./build/ubench_win $c 5 1 2>&1 | grep -E "total best|mismatch|S1|in-window|err"
done
echo "== ncu C"
UBENCH_ONLY="W r8 s3" timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_window_count -s 2 -c 1 -o gpurun_out/r2_prof_window_03 -f ./build/ubench_win C 1 0 2>&1 | tail -3
} > $out 2>&1
cat $out
