#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page raw --csv` (one kernel launch) into the handful of counters DESIGN.md quotes.
Usage: python tools/ncu_summary.py raw.csv [source.csv] > profiles/rNN_ncu_<kernel>_summary.txt"""
import csv
import subprocess
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[-1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
print("# source: %s (ncu --set full --clock-control none, one launch inside `bench.py --steps 1 --warmup 3`)" % sys.argv[1])
for w in want:
    for h, u, v in zip(hdr, units, vals):
        if h == w:
            print("%-90s %-12s %s" % (h, u, v))
if len(sys.argv) > 2:
    print()
    print("# hottest source lines (warp stall samples)")
    print(subprocess.run([sys.executable, "tools/ncu_hot_lines.py", sys.argv[2], "25"], capture_output=True, text=True).stdout)
