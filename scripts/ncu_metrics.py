#!/usr/bin/env python
"""Prints selected metrics per kernel launch from an .ncu-rep (read with `ncu -i ... --page raw --csv`).
Usage: ncu_metrics.py report.ncu-rep [extra_metric_substring ...]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active",
        "lts__t_sectors_srcunit_tex.sum", "smsp__cycles_active.avg"]
extra = sys.argv[2:]
for r in rows[2:]:
    print("---")
    for i, h in enumerate(hdr):
        if h in want or any(e in h for e in extra):
            print("  %-70s %s %s" % (h, r[i][:60], units[i]))
