#!/usr/bin/env python3
"""gpurun_out/traffic_<tag>.csv (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum at the
bench's full 1 GiB size) -> profiles/<tag>_traffic.{csv,json}; bench.py reads the json for roofline.traffic."""
import collections, csv, json, os, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_hash   # the capture is only quoted by bench.py for the kernel sources it was taken from
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
src = os.path.join(ROOT, "gpurun_out", "traffic_%s.csv" % tag)
rows = list(csv.reader(open(src)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
col = {n: i for i, n in enumerate(rows[h])}
per = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) <= col["Metric Value"]:
        continue
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("mh::<unnamed>::", "")
    per.setdefault((r[col["ID"]], name), {})[r[col["Metric Name"]]] = (float(r[col["Metric Value"]].replace(",", "")), r[col["Metric Unit"]])
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
last = {}
for (_, name), m in per.items():
    last[name] = m
out = {}
for name, m in last.items():
    rd = m["dram__bytes_read.sum"][0] * scale.get(m["dram__bytes_read.sum"][1], 1)
    wr = m["dram__bytes_write.sum"][0] * scale.get(m["dram__bytes_write.sum"][1], 1)
    out[name] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "traffic_bytes": rd + wr, "ncu_duration": list(m["gpu__time_duration.sum"])}
    print(name, out[name])
json.dump({"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none; python bench.py --steps 1 --warmup 3 "
                     "--config markov --no-cpu-baseline --no-e2e (1 GiB Markov text); last launch of each kernel", "input_bytes": 1 << 30, "config": "markov",
           "source_hash": kernel_source_hash(), "kernels": out},
          open(os.path.join(ROOT, "profiles", "%s_traffic.json" % tag), "w"), indent=1)
shutil.copyfile(src, os.path.join(ROOT, "profiles", "%s_traffic.csv" % tag))
