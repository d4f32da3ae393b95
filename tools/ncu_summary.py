#!/usr/bin/env python
"""Per-kernel summary (one row per profiled launch) of an `ncu --set full` report -> CSV for profiles/.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/ncu_full_summary.csv"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__cycles_active.avg",
        "sm__cycles_elapsed.avg"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keys = [k for k in KEYS if k in hdr]
w = csv.writer(sys.stdout)
w.writerow(["Kernel Name"] + keys)
w.writerow([""] + [units[hdr.index(k)] for k in keys])
for r in rows[2:]:
    w.writerow([r[hdr.index("Kernel Name")]] + [r[hdr.index(k)] for k in keys])
