#!/usr/bin/env python3
"""Static SASS instruction histogram per kernel (development aid).
usage: sasscount.py <binary-or-so> [name-substring]"""
import collections, re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
flt = sys.argv[2] if len(sys.argv) > 2 else ""
cur = None
hist = collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
for k, c in hist.items():
    if flt in k:
        tot = sum(c.values())
        print(f"{k}: {tot} instrs")
        print("   " + ", ".join(f"{op}:{n}" for op, n in c.most_common(18)))
