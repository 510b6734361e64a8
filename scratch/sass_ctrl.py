#!/usr/bin/env python
"""Decode the scheduling control fields of sm_100 SASS (cuobjdump -sass output on stdin or a file): for every
instruction print its write / read scoreboard, the scoreboards it waits for, and its stall count. Used to find loads
that share a scoreboard with the prefetch issued behind them (the first use of the OLD data then waits for the NEW loads).
usage: cuobjdump -sass -fun <kernel> lib.so | python scratch/sass_ctrl.py [pattern]"""
import re, sys
pat = re.compile(sys.argv[1]) if len(sys.argv) > 1 else None
lines = sys.stdin.read().splitlines()
ins = re.compile(r'^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/')
hi = re.compile(r'^\s+/\* 0x([0-9a-f]{16}) \*/')
out = []
i = 0
while i < len(lines):
    m = ins.match(lines[i])
    if m and i + 1 < len(lines):
        h = hi.match(lines[i + 1])
        if h:
            c = int(h.group(1), 16) >> 41
            stall, yld, wr, rd, wait = c & 15, (c >> 4) & 1, (c >> 5) & 7, (c >> 8) & 7, (c >> 11) & 63
            out.append((m.group(1), m.group(2).strip(), stall, wr, rd, wait))
            i += 2
            continue
    i += 1
for a, t, stall, wr, rd, wait in out:
    s = "%s  %-70s st=%2d wr=%s rd=%s wait=%s" % (a, t[:70], stall, wr if wr != 7 else '-', rd if rd != 7 else '-',
                                                   ''.join(str(b) for b in range(6) if wait >> b & 1) or '-')
    if pat is None or pat.search(t) or wait:
        print(s)
