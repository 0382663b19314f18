#!/usr/bin/env python
"""Print the kernels of one steady-state step from an ncu launch list (gpu__time_duration).  usage: launch_step.py <csv>"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[h]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
seq = [(re.sub(r'void hrb::\(anonymous namespace\)::|\(hrb::SearchArgs\)', '', r[ki])[:64], float(r[vi].replace(',', ''))) for r in rows[h + 1:] if len(r) > vi]
idx = [i for i, (k, _) in enumerate(seq) if 'packFrame' in k]
a, b = idx[1], idx[2]
tot = 0
for k, v in seq[a:b]:
    print(f'{v / 1000:8.1f} {k}'); tot += v
print(f'{tot / 1000:8.1f} total')
