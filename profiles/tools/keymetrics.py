#!/usr/bin/env python
"""Print the handful of ncu metrics the step-kernel summaries quote.  usage: keymetrics.py <report.ncu-rep>"""
import csv, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
r = list(csv.reader(txt.splitlines()))
h, v = r[0], r[-1]
keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__grid_size",
        "launch__block_size", "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "launch__waves_per_multiprocessor",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled", "smsp__average_warp_latency_issue_stalled"]
for k in keys:
    for i, x in enumerate(h):
        if x.startswith(k) and "per_second" not in x and "pct_of_peak_sustained_elapsed" not in x:
            print("%-90s %s" % (x, v[i]))
