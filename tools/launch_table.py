#!/usr/bin/env python
"""Per-launch table from an ncu --csv launch list: python tools/launch_table.py file.csv [max_rows]"""
import csv, sys
from collections import OrderedDict
rows = list(csv.reader(open(sys.argv[1])))
s = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
h = rows[s]; d = OrderedDict()
for r in rows[s + 1:]:
    if len(r) < len(h): continue
    k = r[h.index('ID')]
    d.setdefault(k, {'name': r[h.index('Kernel Name')][:44], 'grid': r[h.index('Grid Size')], 'block': r[h.index('Block Size')]})
    d[k][r[h.index('Metric Name')]] = r[h.index('Metric Value')]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
for k, v in list(d.items())[:n]:
    print(k, v['name'], v['grid'], v['block'], 'ns=' + str(v.get('gpu__time_duration.sum')), 'inst=' + str(v.get('smsp__inst_executed.sum')), 'regs=' + str(v.get('launch__registers_per_thread')))
