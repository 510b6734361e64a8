"""Executed warp instructions of one ncu capture by code region: runs of consecutive SASS instructions with the same
execution count (one line per run that holds at least 0.4 % of the kernel's instructions), with the average number of
active threads — shows which divergent paths a warp steps through and what they cost.   usage: inst_regions.py X.ncu-rep"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; col = {n:i for i,n in enumerate(hdr)}
tot = 0; data=[]
for r in rows[2:]:
    if len(r) < len(hdr): continue
    ie = int(r[col["Instructions Executed"]] or 0); te = int(r[col["Thread Instructions Executed"]] or 0)
    data.append((r[col["Address"]], r[col["Source"]].strip(), ie, te, int(r[col["# Samples"]] or 0)))
    tot += ie
print("total warp inst", tot)
# print compressed listing: group consecutive instructions with same exec count
i=0
while i < len(data):
    j=i
    while j+1 < len(data) and data[j+1][2]==data[i][2]: j+=1
    n=j-i+1
    if data[i][2]*n > tot*0.004:
        print("%6.2f%%  n=%3d exec=%9d avgthr=%4.1f  %s .. %s" % (100.0*data[i][2]*n/tot, n, data[i][2], (sum(d[3] for d in data[i:j+1])/max(1,sum(d[2] for d in data[i:j+1]))), data[i][1][:40], data[j][1][:40]))
    i=j+1
