#!/usr/bin/env python3
"""Per-source-line stall samples of an ncu report: python profiles/src_stalls.py X.ncu-rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[h]
col = {}
for i, n in enumerate(hdr):
    col.setdefault(n, i)
stall_cols = [i for i, n in enumerate(hdr) if n.startswith("stall_")]
seen = set(); sc = []
for i in stall_cols:
    if hdr[i] not in seen:
        seen.add(hdr[i]); sc.append(i)
lines = []
for r in rows[h + 1:]:
    if len(r) < len(hdr) or not r[0]:
        continue   # SASS rows have an empty line number
    try:
        s = int(r[col["# Samples"]] or 0)
    except ValueError:
        continue
    per = {hdr[i]: int(r[i] or 0) for i in sc if (r[i] or "0") != "0"}
    lines.append((s, int(r[0]), r[1].strip()[:100], r[col["Instructions Executed"]], per))
tot = sum(l[0] for l in lines)
print("total samples", tot)
for s, ln, src, ex, per in sorted(lines, key=lambda t: -t[0])[:top]:
    t = ", ".join("%s=%d" % (k.replace("stall_", ""), v) for k, v in sorted(per.items(), key=lambda kv: -kv[1])[:3])
    print("%6d %5.1f%%  L%-4d exec=%-9s %s   [%s]" % (s, 100.0 * s / max(1, tot), ln, ex, src, t))
