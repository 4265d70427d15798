#!/usr/bin/env python3
"""One CSV row per kernel of an `ncu --set full` report with the columns the summaries quote.
usage: ncu_extract.py report.ncu-rep > metrics.csv"""
import csv
import subprocess
import sys

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "smsp__inst_executed.sum", "sm__cycles_active.avg", "sm__cycles_active.max",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum"]
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
keep = ["Kernel Name", "Block Size", "Grid Size"] + [c for c in COLS if c in idx]
w = csv.writer(sys.stdout)
w.writerow(keep)
w.writerow([units[idx[c]] for c in keep])
for r in rows[2:]:
    w.writerow([r[idx[c]] for c in keep])
