#!/usr/bin/env python
"""Instruction counts per kernel from `cuobjdump -sass`: total, BRA.DIV, uniform-datapath, R2UR.
  python tools/sass_count.py [substring] [lib]"""
import re
import subprocess
import sys

pat = sys.argv[1] if len(sys.argv) > 1 else "k_sweep_n16"
lib = sys.argv[2] if len(sys.argv) > 2 else "lattice_qft_b200/liblqft.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
name, rows = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        rows[name] = [0, 0, 0, 0]
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(\S+)\s+(\S+)", line)
    if m and name:
        op = m.group(2) if m.group(1).startswith("@") else m.group(1)
        r = rows[name]
        r[0] += 1
        r[1] += op.startswith("BRA.DIV")
        r[2] += op.startswith("U") or "UR" in line.split("/*")[1]
        r[3] += op.startswith("R2UR")
demangle = subprocess.run(["c++filt"], input="\n".join(rows), capture_output=True, text=True).stdout.splitlines()
for n, d in zip(rows, demangle):
    if pat in d:
        t = rows[n]
        print(f"{t[0]:6d} instr  bra.div={t[1]:3d}  uniform={t[2]:4d}  r2ur={t[3]:3d}  {d.split('(')[0]}")
