#!/usr/bin/env python
"""Condense ncu CSV exports (gpurun_out/prof/*) into the tracked summaries under profiles/.
usage: python tools/ncu_summary.py <tag>"""
import collections
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def short(name):
    return name.replace("void ", "").replace("unnamed>::", "").replace("hrb::", "").split("(")[0]


def launches(tag, out):
    path = os.path.join(ROOT, "gpurun_out", "prof", f"launches_{tag}.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    hdr, data = None, []
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    agg = collections.OrderedDict()
    for d in data:
        a = agg.setdefault(short(d["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += float(d["Metric Value"]) / 1e3
    bench = {k: v for k, v in agg.items() if "sadPeak" not in k}
    tot = sum(v[1] for v in bench.values())
    out.write(f"## launch list ({len(data)} launches, `ncu --metrics gpu__time_duration.sum --clock-control none`, cold-cache serialised: compare shares)\n\n")
    out.write("| kernel | launches | total us | avg us | share of step kernels |\n|---|---|---|---|---|\n")
    for k, v in sorted(bench.items(), key=lambda kv: -kv[1][1]):
        out.write(f"| `{k}` | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.1f} | {v[1] / tot * 100:.1f}% |\n")
    out.write("\n")


def raw(tag, which, out, picks=None):
    path = os.path.join(ROOT, "gpurun_out", "prof", f"{which}_raw_{tag}.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    sel = picks if picks is not None else range(len(data))
    sel = [i for i in sel if i < len(data)]
    out.write(f"## `ncu --set full` — {which} ({len(data)} captured launches; shown: {list(sel)})\n\n")
    out.write("| metric | unit | " + " | ".join(f"#{i} {short(data[i][idx['Kernel Name']])[:28]}" for i in sel) + " |\n")
    out.write("|---|---|" + "---|" * len(sel) + "\n")
    for k in KEYS:
        if k not in idx:
            continue
        out.write(f"| {k} | {units[idx[k]]} | " + " | ".join(data[i][idx[k]] for i in sel) + " |\n")
    out.write("\n")


def step(tag, out, workload):
    """Every kernel of one steady-state step from the `--set full` capture; also writes profiles/ncu_traffic.json."""
    import json
    path = os.path.join(ROOT, "gpurun_out", "prof", f"step_raw_{tag}.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    names = [short(d[idx["Kernel Name"]]) for d in data]
    # one step = from a packFrameKernel to the next
    packs = [i for i, n in enumerate(names) if "packPlanar" in n or "packFrame" in n]
    a, b = (packs[0], packs[1]) if len(packs) > 1 else (0, len(data))
    sel = list(range(a, b))

    def val(i, k):
        try:
            return float(data[i][idx[k]].replace(",", ""))
        except Exception:  # noqa: BLE001
            return float("nan")

    def mb(i, k):  # ncu prints bytes with a per-column unit
        u = units[idx[k]].lower()
        scale = {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, 1.0)
        return val(i, k) * scale

    out.write(f"## `ncu --set full --clock-control none` — every kernel of one steady-state step ({len(sel)} launches)\n\n")
    cols = [("us", "gpu__time_duration.sum"), ("regs", "launch__registers_per_thread"), ("grid", "launch__grid_size"),
            ("warps act %", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("issue act %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            ("ALU pipe %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
            ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("L2 hit %", "lts__t_sector_hit_rate.pct"),
            ("warp-instr M", "smsp__inst_executed.sum"),
            ("stall long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
            ("stall math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
            ("stall barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio")]
    out.write("| # | kernel | " + " | ".join(c for c, _ in cols) + " | DRAM rd MB | DRAM wr MB |\n|---|---|" + "---|" * (len(cols) + 2) + "\n")
    traffic = collections.OrderedDict()
    tot_us = 0.0
    for i in sel:
        cells = []
        for c, k in cols:
            v = val(i, k)
            if c == "warp-instr M":
                v /= 1e6
            cells.append(f"{v:.1f}" if c not in ("regs", "grid") else f"{v:.0f}")
        rd, wr = mb(i, "dram__bytes_read.sum"), mb(i, "dram__bytes_write.sum")
        tot_us += val(i, "gpu__time_duration.sum")
        out.write(f"| {i - a} | `{names[i][:48]}` | " + " | ".join(cells) + f" | {rd:.1f} | {wr:.1f} |\n")
        key = names[i].split("<")[0]
        for kk in ([key] + (["search_pass"] if key.startswith("sad") else [])):   # search_pass: every pass kernel of the ladder together
            t = traffic.setdefault(kk, [0, 0.0])
            t[0] += 1
            t[1] += (rd + wr) * 1e6
    out.write(f"\nSum of the step's kernel durations under ncu (cold caches, serialised): {tot_us:.0f} us.\n\n")
    tj = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        allw = json.load(open(tj))
    except Exception:  # noqa: BLE001
        allw = {}
    allw[workload] = {k: {"launches": v[0], "dram_bytes_per_launch": int(v[1] / v[0]), "source": f"profiles/ncu_summary_{tag}.md"} for k, v in traffic.items()}
    json.dump(allw, open(tj, "w"), indent=1)
    out.write("Per-kernel DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum, mean per launch) is also written to `profiles/ncu_traffic.json`, "
              "which `bench.py` reports as `roofline.traffic`.\n\n")


def main():
    tag = sys.argv[1]
    workload = sys.argv[2] if len(sys.argv) > 2 else "cfg3"
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    dst = os.path.join(ROOT, "profiles", f"ncu_summary_{tag}.md")
    with open(dst, "w") as out:
        out.write(f"# ncu summary {tag} ({workload})\n\nCommands: `python bench.py --workload {workload} --steps 2 --warmup 3 --no-cpu-baseline` (launch list) and `... --steps 4 --warmup 6 ...` with the first seven steps skipped (`--set full` capture of one steady-state step) under ncu "
                  "(tools/gpu_profile.sh); numbers taken under the profiler are for SHARES and per-kernel counters only, never bench values.\n\n")
        launches(tag, out)
        step(tag, out, workload)
    print("wrote", dst)


if __name__ == "__main__":
    main()
