#!/bin/bash
# Kernel-A variants that were prepared but not (fully) measured in round 1 -- see DESIGN.md section 5.
# Build them HERE first (nvcc cross-compiles without a GPU), then run this script on the GPU box:
#     make experiments && gpurun --timeout 300 -- 'bash tools/run_experiments.sh C'
# Every variant re-checks its result against the oracle (PARITY OK) and, for the packed counters,
# the decoded total against the number of steps (no field overflowed).
cfg=${1:-C}
F="deferred-OR <5> grid 10x|PACKED|PARITY|max depth|rror"
for b in ubench ubench_p16 ubench_p16m ubench_p8 ubench_p8m; do
    [ -x build/$b ] || { echo "build/$b missing: run 'make experiments' first"; continue; }
    echo "== $b"
    timeout 60 env UBENCH_MERGED=1 UBENCH_ONLY_MERGED=1 ./build/$b "$cfg" 6 1 2>&1 | grep -E "$F"
done
