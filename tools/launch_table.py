#!/usr/bin/env python
"""Per-kernel table of an ncu launch list (gpu__time_duration.sum): usage: python tools/launch_table.py <csv> [--seq]"""
import collections, csv, sys

def short(n):
    return n.replace("void ", "").replace("<unnamed>::", "").replace("hrb::", "").split("(hrb")[0].split("(const")[0]

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = None, []
for r in rows:
    if r and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(dict(zip(hdr, r)))
if "--seq" in sys.argv:
    for d in data:
        print(f"{float(d['Metric Value'])/1e3:9.2f} us  {short(d['Kernel Name'])}")
agg = collections.OrderedDict()
for d in data:
    a = agg.setdefault(short(d["Kernel Name"]), [0, 0.0])
    a[0] += 1
    a[1] += float(d["Metric Value"]) / 1e3
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:10.1f} us total  {v[0]:4d} x {v[1]/v[0]:8.2f} us  {100*v[1]/tot:5.1f}%  {k}")
print(f"{tot:10.1f} us total over {len(data)} launches")
