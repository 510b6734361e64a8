#!/usr/bin/env python3
"""Print the SASS around the hottest line(s) of an ncu report: python scratch/sass_ctx.py X.ncu-rep [before] [after] [rank]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; before = int(sys.argv[2]) if len(sys.argv) > 2 else 40; after = int(sys.argv[3]) if len(sys.argv) > 3 else 6
rank = int(sys.argv[4]) if len(sys.argv) > 4 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
col = {n: i for i, n in enumerate(rows[h])}
body = [r for r in rows[h + 1:] if len(r) >= len(rows[h])]
order = sorted(range(len(body)), key=lambda i: -int(body[i][col["# Samples"]] or 0))
b = order[rank]
for i in range(max(0, b - before), min(len(body), b + after)):
    r = body[i]
    print("%5d %6s %9s  %s" % (i, r[col["# Samples"]], r[col["Instructions Executed"]], r[col["Source"]]))
