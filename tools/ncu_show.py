#!/usr/bin/env python
"""Print key metrics of every launch in an ncu raw CSV export.  usage: python tools/ncu_show.py <raw.csv> [extra-metric-substrings]"""
import csv, sys
KEYS = ["gpu__time_duration.sum","launch__registers_per_thread","launch__occupancy_limit_registers","launch__occupancy_limit_shared_mem","sm__warps_active.avg.pct_of_peak_sustained_active",
"smsp__issue_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active","sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
"l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed","l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed","l1tex__t_sector_hit_rate.pct","lts__t_sector_hit_rate.pct","lts__t_sectors.avg.pct_of_peak_sustained_elapsed","lts__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
"gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","dram__bytes_read.sum","dram__bytes_write.sum","lts__t_bytes.sum","l1tex__m_xbar2l1tex_read_bytes.sum","smsp__inst_executed.sum","l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum","l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","l1tex__data_pipe_lsu_wavefronts.sum",
"smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio","smsp__average_warps_issue_stalled_wait_per_issue_active.ratio","smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio","smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio","smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio","smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
extra = [h for h in hdr for sub in sys.argv[2:] if sub in h]
names = [d[idx["Kernel Name"]].replace("void unnamed>::", "").split("(")[0][:26] for d in data]
print(f"{'metric':100s} " + " | ".join(f"{n:>16s}" for n in names))
for k in KEYS + extra:
    if k in idx:
        print(f"{k[:92]:92s} {units[idx[k]][:7]:7s} " + " | ".join(f"{d[idx[k]][:16]:>16s}" for d in data))
