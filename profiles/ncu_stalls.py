#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv` dump: total stall samples by reason and the hottest SASS lines.
usage: ncu -i X.ncu-rep --page source --csv | python profiles/ncu_stalls.py [top_n]"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
# the first line is the kernel name record; find the header row
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
col = {name: i for i, name in enumerate(hdr)}
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
top_n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
tot = {n: 0 for n in stall_cols}
lines = []
for r in rows[h + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        samples = int(r[col["# Samples"]] or 0)
    except ValueError:
        continue
    per = {}
    for n in stall_cols:
        try:
            v = int(r[col[n]] or 0)
        except ValueError:
            v = 0
        tot[n] += v
        if v:
            per[n] = v
    lines.append((samples, r[col["Address"]], r[col["Source"]], r[col["Instructions Executed"]], per))
allsamp = sum(s for s, *_ in lines)
print("total samples", allsamp)
for n, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v:
        print("  %-28s %8d  %5.1f%%" % (n, v, 100.0 * v / max(1, allsamp)))
print("hottest instructions:")
for s, addr, src, execd, per in sorted(lines, key=lambda t: -t[0])[:top_n]:
    top = ", ".join("%s=%d" % (k.replace("stall_", ""), v) for k, v in sorted(per.items(), key=lambda kv: -kv[1])[:3])
    print("  %6d %5.1f%%  %-70s exec=%s  [%s]" % (s, 100.0 * s / max(1, allsamp), src[:70], execd, top))
