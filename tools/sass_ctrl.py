#!/usr/bin/env python3
"""Decode the scheduling control field of sm_100 SASS (cuobjdump -sass output): for every instruction print
stall count, write-barrier, read-barrier and wait mask, so the scoreboard a load signals and the
instruction that waits on it can be read off.  usage: sass_ctrl.py file.sass [regex-filter]"""
import re, sys
lines = open(sys.argv[1]).read().split('\n')
flt = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
ins = re.compile(r'^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/')
hi = re.compile(r'^\s*/\* (0x[0-9a-f]{16}) \*/')
i = 0
while i < len(lines):
    m = ins.match(lines[i])
    if m and i + 1 < len(lines):
        h = hi.match(lines[i + 1])
        if h:
            w = int(h.group(1), 16)
            ctrl = w >> 41
            stall = ctrl & 0xF; yld = (ctrl >> 4) & 1; wr = (ctrl >> 5) & 7; rd = (ctrl >> 8) & 7; wait = (ctrl >> 11) & 0x3F
            txt = m.group(2).strip()
            if not flt or flt.search(txt) or wait:
                print(f"{m.group(1)} st={stall:2d} wr={'-' if wr==7 else wr} rd={'-' if rd==7 else rd} wait={wait:06b}  {txt}")
            i += 2
            continue
    i += 1
