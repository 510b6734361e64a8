#!/usr/bin/env python3
"""Turn the ncu outputs brought back in gpurun_out/ into small text summaries that can be committed.
usage: python profiles/summarize.py <tag>      (reads gpurun_out/<tag>_<kernel>.ncu-rep and gpurun_out/launches_<tag>.csv)"""
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def raw_metrics(rep):
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    name = vals[col["Kernel Name"]] if "Kernel Name" in col else "?"
    out = ["kernel: " + name]
    for m in METRICS:
        if m in col:
            out.append("%-70s %18s %s" % (m, vals[col[m]], units[col[m]]))
    return "\n".join(out)


def stalls(rep, top=14):
    src = ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"])
    p = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "ncu_stalls.py"), str(top)], input=src, stdout=subprocess.PIPE, text=True)
    return p.stdout


def src_stalls(rep, top=14):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "src_stalls.py"), rep, str(top)], stdout=subprocess.PIPE, text=True)
    return p.stdout


def regions(rep):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "inst_regions.py"), rep], stdout=subprocess.PIPE, text=True)
    return p.stdout


def launch_list(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    col = {n: i for i, n in enumerate(rows[h])}
    agg = {}
    for r in rows[h + 1:]:
        if len(r) <= col["Metric Value"] or r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("mh::<unnamed>::", "")
        v = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]]
        us = v / 1000.0 if unit in ("ns", "nsecond") else (v * 1000.0 if unit in ("ms", "msecond") else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += us
    tot = sum(a[1] for a in agg.values())
    out = ["%-60s %8s %12s %7s" % ("kernel", "launches", "total us", "share")]
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%-60s %8d %12.1f %6.1f%%" % (k[:60], n, us, 100 * us / tot))
    return "\n".join(out)


def main():
    ll = os.path.join(OUT, "launches_%s.csv" % tag)
    if os.path.exists(ll):
        with open(os.path.join(ROOT, "profiles", "%s_launches.txt" % tag), "w") as fh:
            fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)\n")
            fh.write("# command: python bench.py --config markov --steps 1 --warmup 3 --no-cpu-baseline --no-e2e (1 GiB)\n")
            fh.write(launch_list(ll) + "\n")
    for k in ("hist_lane_kernel", "hist_kernel", "tables_build_kernel", "encode_kernel", "dec_sync_kernel", "dec_write_kernel"):
        rep = os.path.join(OUT, "%s_%s.ncu-rep" % (tag, k))
        if not os.path.exists(rep):
            continue
        with open(os.path.join(ROOT, "profiles", "%s_%s.txt" % (tag, k)), "w") as fh:
            fh.write("# ncu --set full --clock-control none --import-source on, 1 launch after warm-up, the bench workload (1 GiB Markov text)\n")
            fh.write(raw_metrics(rep) + "\n\n# warp stall samples by SASS instruction\n" + stalls(rep))
            fh.write("\n# warp stall samples by source line (-lineinfo)\n" + src_stalls(rep))
            fh.write("\n# executed warp instructions by code region (runs of SASS instructions with the same execution count)\n" + regions(rep))
        print("wrote", k)


if __name__ == "__main__":
    main()
