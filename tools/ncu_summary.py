#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into the handful of counters the roofline needs.
usage: ncu_summary.py <file.ncu-rep> [more.ncu-rep ...]"""
import csv, io, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "lts__t_requests_op_red.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__shared_mem_per_block_static", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]
for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        rec = dict(zip(hdr, vals))
        print(f"== {path}: {rec.get('Kernel Name')}  grid {rec.get('Grid Size')} block {rec.get('Block Size')}")
        for h, u, v in zip(hdr, units, vals):
            if h in KEYS or ("warp_issue_stalled" in h and h.endswith("per_warp_active.pct")):
                try:
                    if float(v.replace(",", "")) == 0.0 and "stalled" in h:
                        continue
                except ValueError:
                    pass
                print(f"  {h:78s} {v:>18s} {u}")
