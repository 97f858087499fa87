#!/usr/bin/env python3
"""Wall-clock of the `fgfa` CLI on a synthetic .flatgfa file, the way the reference's own
harness times it (bench/bench.py:68-85: warm-up + repeated runs of
`fgfa -i X.flatgfa depth`, bench/config.toml:29-32), for node depth (-d) and path depth.
usage: python tools/bench_cli.py [B|C] [runs]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pollen_b200 import flatgfa_io, synth  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "B"
    runs = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    cfg = synth.CONFIGS[name]
    steps, s, e = synth.make_graph(cfg)
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    path = os.path.join(ROOT, "build", f"synth_{name}.flatgfa")
    flatgfa_io.write_flatgfa(path, steps, s, e, cfg.n_segs)
    fgfa = os.path.join(ROOT, "bin", "fgfa")
    out = {"config": name, "file_bytes": os.path.getsize(path), "runs": runs}
    for mode, args in (("node_depth", ["depth", "-d"]), ("path_depth", ["depth"])):
        times = []
        for i in range(runs + 1):                      # first run = warm-up (page cache, CUDA init)
            t0 = time.perf_counter()
            r = subprocess.run([fgfa, "-i", path] + args, stdout=subprocess.PIPE, check=True)
            dt = time.perf_counter() - t0
            if i:
                times.append(dt)
        out[mode] = {"best_s": min(times), "mean_s": sum(times) / len(times), "stdout_bytes": len(r.stdout)}
    print(json.dumps(out))
    os.remove(path)


if __name__ == "__main__":
    main()
