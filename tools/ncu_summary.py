#!/usr/bin/env python
"""Print the handful of ncu raw-page metrics we track from a .ncu-rep (run where ncu is installed)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_fp64.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_lsu.sum",
        "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_uniform.sum",
        "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fp64.sum",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum"]
for i, h in enumerate(hdr):
    if h in keys or ("issue_stalled" in h and "per_issue_active" in h):
        vals = [r[i] for r in data]
        if "issue_stalled" in h and all(float(v) < 0.05 for v in vals):
            continue
        print(f"{h} [{units[i]}] {vals}")
