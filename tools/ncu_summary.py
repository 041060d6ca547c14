#!/usr/bin/env python3
"""prints the handful of ncu metrics this project tracks from a .ncu-rep (run where ncu is installed):
python tools/ncu_summary.py gpurun_out/tile_c4.ncu-rep"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio"]

rows = list(csv.reader(subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
    for i, h in enumerate(hdr):
        if h in WANT or ("issue_stalled" in h and h.endswith("per_warp_active.pct")):
            try:
                v = float(vals[i].replace(",", ""))
            except ValueError:
                continue
            if "issue_stalled" in h and v < 3.0:
                continue
            print(f"  {h} [{units[i]}] = {vals[i]}")
