"""Exercises the host-only code paths (printers, BED parser, windows, block table formatter, threaded\nparser) -- meant to be run under the sanitizers: see tools/asan_host.sh."""
import os, sys, glob
import numpy as np
import pollen_b200 as pb
from pollen_b200 import flatgfa_py, flash
# exercise the remaining host code under the sanitizers: printers, BED parser, windows, big table formatter, threaded parser
for f in sorted(glob.glob("tests/golden/*.gfa")):
    g = flatgfa_py.parse(f)
    assert str(g).encode() == g._h.format_gfa()
    g._h.image()
for text in (b"x\t0\t4\nx\t4\t8\n", b"#c\nx\t1 2\n", b"x\t\n", b"", b"x\t4\t\n", b"\t1\t2\nq\t18446744073709551615\t99999999999999999999999\n"):
    try:
        pb.FlatBED.parse(text).entries()
    except pb.DepthError:
        pass
pb.FlatBED.windows(b"p", 3, 1000, 7).entries()
from pollen_b200 import flatgfa_io
n = (1 << 18) + 77
img = flatgfa_io.build_image(np.zeros(1, np.uint32), [0], [1], n)
img.tofile("build/t_smoke.flatgfa")
with pb.FlatGFA.load("build/t_smoke.flatgfa") as g:
    d = np.arange(n, dtype=np.uint64) * np.uint64(1 << 40)
    assert len(g.format_seg_depth(d, d)) > n
os.remove("build/t_smoke.flatgfa")
rng = np.random.default_rng(1)
toks = b",".join(b"%d%s" % (int(x), (b"+", b"-")[int(o)]) for x, o in zip(rng.integers(1, 50, 900_000), rng.integers(0, 2, 900_000)))
text = b"".join(b"S\t%d\tAC\n" % i for i in range(1, 50)) + b"P\tgiant\t" + toks + b"\t*\nP\tq\t1+\t*\n"
g = pb.FlatGFA.parse_bytes(text)
assert g.path_step_count(0) == 900_000
g.close()
print("sanitized host paths ok")
