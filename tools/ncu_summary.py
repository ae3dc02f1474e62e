"""Summarise an .ncu-rep (raw page) into the few metrics we track. usage: ncu_summary.py rep [out.txt]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'lts__t_bytes.sum','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread',
 'launch__grid_size','launch__waves_per_multiprocessor','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','smsp__inst_executed.sum',
 'smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_lsu.sum','smsp__inst_executed_op_shared_ld.sum','smsp__inst_executed_op_shared_st.sum','smsp__inst_executed_op_global_ld.sum','smsp__inst_executed_op_global_st.sum',
 'smsp__inst_executed_op_local_ld.sum','smsp__inst_executed_op_local_st.sum',
 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio',
 'smsp__warps_eligible.avg.per_cycle_active','smsp__warps_active.avg.per_cycle_active']
out = []
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    out.append("== " + name[:110])
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            out.append(f"  {k:88s} {r[i]:>18s} {units[i]}")
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
