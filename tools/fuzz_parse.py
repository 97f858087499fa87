"""Mutation fuzzer for the host GFA parser and writers (crash hunting; run from the repo root)."""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pollen_b200 as pb
random.seed(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
alpha = b"\t\t\t\n\n,,+-+-*0123456789SLPHACGTMNDIxz:# "
seeds = [b"H\tVN:Z:1.0\n", b"S\t1\tACGT\n", b"S\t2\tA\tLN:i:1\n", b"L\t1\t+\t2\t-\t3M\n", b"P\tp\t1+,2-\t*\n", b"P\tq\t2+\t1M,2M\n"]
ok = bad = 0
for it in range(60000):
    parts = []
    for _ in range(random.randrange(1, 7)):
        s = bytearray(random.choice(seeds))
        for _ in range(random.randrange(0, 4)):
            op = random.randrange(3)
            pos = random.randrange(len(s) + 1)
            if op == 0 and s: del s[min(pos, len(s) - 1)]
            elif op == 1: s.insert(pos, random.choice(alpha))
            elif s: s[min(pos, len(s) - 1)] = random.choice(alpha)
        parts.append(bytes(s))
    text = b"".join(parts)
    try:
        g = pb.FlatGFA.parse_bytes(text)
        try:
            g.format_gfa()
        except pb.DepthError:
            pass
        g.image()
        g.close()
        ok += 1
    except pb.DepthError:
        bad += 1
print("parsed", ok, "rejected", bad)
