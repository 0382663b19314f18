#!/usr/bin/env python
"""Static look at a kernel's SASS without a GPU: instruction count and opcode histogram of every kernel in an object
file whose (mangled) name matches a regex.  usage: python tools/sass_hist.py build/kernels_frame.o warpFastKernelItLi2 [top]

With an ncu `--page source --csv` export (gpurun_out/prof/source_*.csv) as first argument it weights every instruction
by its executed count instead: python tools/sass_hist.py gpurun_out/prof/source_x.csv <launch-index> [top]"""
import collections
import csv
import re
import subprocess
import sys


def from_object(path, pattern, top):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    name, hist = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1) if re.search(pattern, m.group(1)) else None
            if name:
                hist[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", line)
        if name and m:
            ops = [t for t in m.group(1).split() if not t.startswith("@")]
            if ops:
                hist[name][ops[0]] += 1
    for k, h in hist.items():
        print(f"{k}: {sum(h.values())} instructions")
        for op, n in h.most_common(top):
            print(f"    {op:28s} {n}")


def from_ncu_source(path, launch, top):
    rows = list(csv.reader(open(path)))
    secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    s, e = secs[launch], secs[launch + 1]
    print(rows[s][1][:120])
    hist, tot = collections.Counter(), 0
    for r in rows[s + 2:e]:
        try:
            n = int(r[5])
        except (ValueError, IndexError):
            continue
        ops = [t for t in r[1].split() if not t.startswith("@")]
        if ops:
            hist[ops[0]] += n
            tot += n
    print(f"{tot} warp-instructions executed")
    for op, n in hist.most_common(top):
        print(f"    {op:28s} {n:>12d} {100.0 * n / tot:5.1f}%")


if __name__ == "__main__":
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    if sys.argv[1].endswith(".csv"):
        from_ncu_source(sys.argv[1], int(sys.argv[2]), top)
    else:
        from_object(sys.argv[1], sys.argv[2], top)
