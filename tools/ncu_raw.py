#!/usr/bin/env python
"""Key metrics of an `ncu --page raw --csv` export, one column per captured launch.  usage: ncu_raw.py raw.csv [...]"""
import csv, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"]
for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    h = rows[0]
    print("==", f)
    for r in rows[2:]:
        d = dict(zip(h, r))
        print("kernel:", d.get("Kernel Name", "?")[:60])
        for k in KEYS:
            if k in d: print("  %-90s %s" % (k, d[k]))
