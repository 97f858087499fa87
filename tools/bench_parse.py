#!/usr/bin/env python3
"""Text-input path: `fgfa -I X.gfa -o X.flatgfa` with the step lists tokenised on the GPU
(FGFA_GPU_PARSE=1) against the host routes (threads / single thread).  usage: bench_parse.py [B]"""
import json, os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pollen_b200 import synth

name = sys.argv[1] if len(sys.argv) > 1 else "B"
cfg = synth.CONFIGS[name]
steps, s, e = synth.make_graph(cfg)
os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
src, out = os.path.join(ROOT, "build", f"{name}.gfa"), os.path.join(ROOT, "build", f"{name}.flatgfa")
with open(src, "wb") as f:
    f.write(b"H\tVN:Z:1.0\n")
    f.write(("\n".join(f"S\t{i}\tA" for i in range(1, cfg.n_segs + 1)) + "\n").encode())
    for p in range(cfg.n_paths):
        h = steps[s[p]:e[p]]
        toks = np.char.add(((h >> 1) + 1).astype(str), np.where(h & 1, "-", "+"))
        f.write(b"P\tp%d\t" % p + ",".join(toks.tolist()).encode() + b"\t*\n")
res = {"config": name, "gfa_bytes": os.path.getsize(src), "n_steps": cfg.n_steps, "host_cores": os.cpu_count()}
ref = None
for label, env in (("gpu_tokenizer", {"FGFA_GPU_PARSE": "1"}), ("host_threads", {"FGFA_GPU_PARSE": "0"})):
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        subprocess.run([os.path.join(ROOT, "bin", "fgfa"), "-I", src, "-o", out], check=True, env={**os.environ, **env})
        best = min(best, time.perf_counter() - t0)
    img = open(out, "rb").read()
    ref = ref or img
    res[label] = {"wall_s": best, "text_MB_per_s": res["gfa_bytes"] / 1e6 / best, "identical_output": img == ref}
# in-process (CUDA context already up): upload + T1 + T2 + download of the steps
import pollen_b200 as pb
text = open(src, "rb").read()
fields, pos = [], 0
for line in text.split(b"\n"):
    if line.startswith(b"P\t"):
        f = line.split(b"\t")
        fields.append((pos + 2 + len(f[1]) + 1, len(f[2])))
    pos += len(line) + 1
pb.tokenize_steps(text, fields, cfg.n_segs, None)
best = 1e9
for _ in range(3):
    t0 = time.perf_counter()
    got, gs, ge = pb.tokenize_steps(text, fields, cfg.n_segs, None)
    best = min(best, time.perf_counter() - t0)
res["gpu_tokenizer_in_process"] = {"wall_s": best, "text_MB_per_s": sum(f[1] for f in fields) / 1e6 / best,
                                   "exact": bool((got == steps).all())}
print(json.dumps(res))
os.remove(src); os.remove(out)
