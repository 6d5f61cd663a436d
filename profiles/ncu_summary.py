#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` into the metric,value,unit table committed under profiles/ (one block per launch).
usage: ncu -i gpurun_out/X.ncu-rep --page raw --csv > /tmp/X.csv; python profiles/ncu_summary.py /tmp/X.csv > profiles/r1_ncu_X.csv"""
import csv, sys
KEEP = ("Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__warps_eligible.avg.per_cycle_active")
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
print("metric,value,unit")
for r in rows[2:]:
    for k in KEEP:
        if k in hdr:
            i = hdr.index(k)
            v = r[i]
            print(f'{k},"{v}",{units[i]}' if "," in v else f"{k},{v},{units[i]}")
    print("-,-,-")
