"""Per-launch headline metrics of an ncu --set full report (one block per kernel launch).
usage: python tools/ncu_table.py X.ncu-rep [kernel-substring]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
K = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
     ("launch__occupancy_limit_registers", "occ_lim_regs"), ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occupancy_pct"),
     ("smsp__inst_executed.sum", "warp_inst"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads_per_inst"),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
     ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
     ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
     ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
     ("lts__t_sectors_srcunit_tex_op_read.sum", "l2_read_sectors_from_l1"),
     ("lts__t_sectors_srcunit_tex_op_atom.sum", "l2_atom_sectors"), ("lts__t_sectors_srcunit_tex_op_red.sum", "l2_red_sectors"),
     ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
     ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
     ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
     ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
     ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
     ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_throttle"),
     ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall_branch"),
     ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_no_inst"),
     ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected"),
     ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio"),
     ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg"),
     ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall_membar"),
     ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "stall_sleeping"),
     ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall_dispatch"),
     ("smsp__average_warps_issue_stalled_drain_per_issue_active.ratio", "stall_drain"),
     ("smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio", "stall_imc")]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    if want not in r[ki]:
        continue
    print("== %s  (launch id %s)" % (r[ki][:70], r[0]))
    for k, label in K:
        if k in hdr:
            i = hdr.index(k)
            print("   %-26s %s %s" % (label, r[i], units[i]))
