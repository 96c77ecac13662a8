#!/usr/bin/env python
"""Print the metrics we track from an .ncu-rep (first kernel matching a substring). Usage: ncu_summary.py rep [substr]"""
import csv, subprocess, sys
rep = sys.argv[1]; sub = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
rows = list(csv.reader(out.splitlines()))
h = rows[0]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    if sub in r[h.index("Kernel Name")]:
        print("kernel:", r[h.index("Kernel Name")])
        for k in KEYS:
            if k in h: print("  %-88s %s %s" % (k, r[h.index(k)], rows[1][h.index(k)]))
        break
