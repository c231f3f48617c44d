#!/usr/bin/env python
"""Summaries of ncu outputs for profiles/: launch-list shares and key raw metrics.

  python tools/ncu_summary.py launches gpurun_out/x_launches.csv
  python tools/ncu_summary.py raw gpurun_out/x.ncu-rep [kernel-substring]
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[start + 1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        name = r[kn].split("(")[0]
        tot[name] += float(r[mv].replace(",", ""))
        cnt[name] += 1
    s = sum(tot.values())
    print(f"# {path}: {sum(cnt.values())} launches, {s/1e3:.1f} us total (ncu: cold-cache, serialised)")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{k:58s} n={cnt[k]:5d} total={v/1e3:10.1f}us avg={v/cnt[k]/1e3:8.2f}us share={100*v/s:5.1f}%")


def raw(path, pat=""):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if pat not in d.get("Kernel Name", ""):
            continue
        print("---", d.get("Kernel Name", "")[:100])
        for k in KEYS:
            if k in d:
                print(f"{k:75s} {d[k]:>18s} {units[hdr.index(k)]}")
        stalls = sorted(((float(v.replace(",", "")), k) for k, v in d.items()
                         if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio")
                         or (k.startswith("smsp__average_warp_latency_issue_stalled") and k.endswith(".ratio"))), reverse=True)
        for v, k in stalls[:8]:
            print(f"  stall {k:80s} {v:8.3f}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        raw(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
