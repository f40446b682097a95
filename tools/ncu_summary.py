#!/usr/bin/env python
"""Print the headline metrics of an .ncu-rep (first kernel) -- used to write profiles/*.md."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "sm__throughput.avg.pct", "launch__occupancy_limit", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__grid_size", "launch__block_size",
        "smsp__average_warps_issue_stalled", "smsp__average_warp_latency_per_inst_issued", "launch__waves", "Kernel Name",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "smsp__warps_eligible.avg.per_cycle_active"]
for vals in rows[2:]:
    print("=" * 100)
    for h, u, v in zip(hdr, units, vals):
        if any(h.startswith(w) for w in want):
            if h.startswith("smsp__average_warps_issue_stalled") and float(v or 0) < 0.3:
                continue
            print(f"{h:90s} {u:12s} {v}")
