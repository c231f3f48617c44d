#!/usr/bin/env python
"""Top stall locations of the first kernel in an ncu report: python tools/ncu_hot.py x.ncu-rep [N]"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ci = {k: i for i, k in enumerate(hdr)}
end = next((i for i in range(2, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))
body = [r for r in rows[2:end] if len(r) == len(hdr)]
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
tot = sum(float(r[ci["# Samples"]] or 0) for r in body)
agg = {k: sum(float(r[ci[k]] or 0) for r in body) for k in stalls}
print("kernel:", rows[0][1][:120])
print("samples:", tot, " by reason:", ", ".join(f"{k[6:]}={100*v/tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
body.sort(key=lambda r: -float(r[ci["# Samples"]] or 0))
for r in body[:n]:
    s = float(r[ci["# Samples"]] or 0)
    top = max(stalls, key=lambda k: float(r[ci[k]] or 0))
    print(f"{100*s/tot:5.1f}%  {r[ci['Source']][:70]:70s} {top[6:]}")
