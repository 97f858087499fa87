#!/usr/bin/env python3
"""Single-process multi-GPU run of config C through the C ABI (fgfa_depth_multi_*), for profilers:
    python tools/multi_peer_probe.py [n_gpus] [peer|nccl] [reps]
ncu can attach to it (one process), e.g. NVLink bytes of kernel X per launch:
    ncu --metrics nvlrx__bytes.sum,nvltx__bytes.sum,gpu__time_duration.sum -k regex:k_uniq_exchange \
        --csv --log-file gpurun_out/x_nvlink.csv python tools/multi_peer_probe.py 8 peer 2
Prints one JSON line: parity against the oracle, wall-clock per query (host-timed; the device-timed
numbers are bench.py's)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pollen_b200 as pb  # noqa: E402
from pollen_b200 import synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else pb.device_count()
    exchange = sys.argv[2] if len(sys.argv) > 2 else "peer"
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    cfg = synth.CONFIGS[os.environ.get("PROBE_CONFIG", "C")]
    steps, s, e = synth.make_graph(cfg)
    m = pb.MultiDepth(list(range(n)), s, e, cfg.n_segs, cfg.n_steps, exchange=exchange)
    owner, dev_steps = m.partition()
    m.upload(steps)
    m.run()
    d, u = m.download()
    ok = None
    if os.environ.get("PROBE_VERIFY", "1") == "1":
        import oracle_lib as O
        rc, od, ou = O.depth_with_uniq(steps, s, e, cfg.n_segs)
        ok = rc == 0 and bool((d == od).all()) and bool((u == ou).all())
    m.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        m.run()
    m.sync()
    dt = (time.perf_counter() - t0) / reps
    print(json.dumps({"config": cfg.name, "n_gpus": n, "exchange": exchange, "parity_vs_oracle": ok,
                      "resident_ms_per_query_host_clock": dt * 1e3, "device_steps": [int(x) for x in dev_steps]}))
    m.close()


if __name__ == "__main__":
    main()
