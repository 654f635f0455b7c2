"""Prints the handful of ncu raw-page metrics quoted in profiles/*.md from an .ncu-rep (run where ncu is installed)."""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "launch__shared_mem_per_block", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "sm__inst_executed.sum.per_cycle_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_xu.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__thread_inst_executed_per_inst_executed.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sass__inst_executed_shared_loads",
        "sass__inst_executed_shared_stores", "sass__inst_executed_global_loads", "sass__inst_executed_global_stores",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "sm__cycles_elapsed.avg", "sm__cycles_active.avg"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
for r in rows[2:]:
    print("#", r[h.index("Kernel Name")][:90])
    for k in KEYS:
        if k in h:
            print(f"{k:88s} {r[h.index(k)]:>16s} {units[h.index(k)]}")
