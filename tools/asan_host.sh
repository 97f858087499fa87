#!/bin/bash
# Host code under AddressSanitizer + UndefinedBehaviorSanitizer (no GPU needed): the two fuzzers and
# the host-path smoke script against build/asan/libflatgfa_asan.so (`make asan`).
set -e
cd "$(dirname "$0")/.."
make asan
export LD_PRELOAD="$(gcc -print-file-name=libasan.so.8 2>/dev/null || echo /usr/lib/x86_64-linux-gnu/libasan.so.8) /usr/lib/x86_64-linux-gnu/libstdc++.so.6"
[ -e "${LD_PRELOAD%% *}" ] || export LD_PRELOAD="/usr/lib/x86_64-linux-gnu/libasan.so.8 /usr/lib/x86_64-linux-gnu/libstdc++.so.6"
export ASAN_OPTIONS=detect_leaks=0
L=build/asan/libflatgfa_asan.so
python tools/run_with_lib.py $L tools/fuzz_parse.py 4
python tools/run_with_lib.py $L tools/fuzz_view.py 2
python tools/run_with_lib.py $L tools/host_paths_smoke.py
