#!/usr/bin/env python3
"""Filters `ncu --page raw --csv` (stdin) down to the metrics cited in profiles/: one 'metric,unit,value' row each."""
import csv
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "sm__inst_executed.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "local_load", "local_store",
        "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled")
rows = list(csv.reader(sys.stdin))
hdr = next((i for i, r in enumerate(rows) if "ID" in r and "Kernel Name" in r), None)
if hdr is None:
    sys.exit("no ncu raw table on stdin")
names, units = rows[hdr], rows[hdr + 1]
w = csv.writer(sys.stdout)
for data in rows[hdr + 2:]:
    if len(data) != len(names):
        continue
    w.writerow(["# kernel", data[names.index("Kernel Name")][:120], "id " + data[names.index("ID")]])
    for n, u, v in zip(names, units, data):
        if any(n.startswith(k) for k in KEEP):
            w.writerow([n, u, v])
