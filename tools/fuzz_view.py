"""Fuzzer for the .flatgfa viewer and the flatgfa-c accessors on corrupted images (run from the repo root)."""
import random, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pollen_b200 as pb
random.seed(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
g = pb.FlatGFA.parse("tests/golden/ref_tiny.gfa")
img = g.image().copy(); g.close()
path = "build/fuzz.flatgfa"
ok = bad = 0
for it in range(20000):
    m = img.copy()
    for _ in range(random.randrange(1, 4)):
        if random.random() < 0.7:
            off = random.randrange(0, 184)              # table of contents
        else:
            off = random.randrange(184, m.size)        # pool contents (spans, handles, ...)
        m[off] = random.choice([0, 1, 2, 3, 7, 8, 16, 255, random.randrange(256)])
    if random.random() < 0.1:
        m = m[: random.randrange(0, m.size)]
    m.tofile(path)
    try:
        h = pb.FlatGFA.load(path)
    except pb.DepthError:
        bad += 1
        continue
    try:
        n = h.segment_count
        for i in range(min(n, 6)): h.seq(i)
        for p in range(min(h.path_count, 4)):
            h.path_name(p); c = h.path_step_count(p)
            if c != 0xFFFFFFFF:
                for s in range(min(c, 5)): h.step(p, s)
        try:
            h.format_gfa()
        except pb.DepthError:
            pass
        ok += 1
    finally:
        h.close()
os.remove(path)
print("viewed", ok, "rejected", bad)
