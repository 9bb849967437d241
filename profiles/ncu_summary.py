import csv, sys, subprocess
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]
want=["Kernel Name","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","sm__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__t_sector_hit_rate.pct","lts__t_sector_hit_rate.pct","smsp__inst_executed.sum","smsp__thread_inst_executed_per_inst_executed.ratio","smsp__issue_active.avg.pct_of_peak_sustained_active","smsp__warps_eligible.avg.per_cycle_active","launch__occupancy_limit_registers","launch__occupancy_limit_shared_mem","launch__shared_mem_per_block_dynamic","sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_alu.sum","sm__inst_executed_pipe_fma.sum","sm__inst_executed_pipe_lsu.sum","sm__inst_executed_pipe_xu.sum","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","smsp__pcsamp_warps_issue_stalled_long_scoreboard","smsp__pcsamp_warps_issue_stalled_short_scoreboard","smsp__pcsamp_warps_issue_stalled_wait","smsp__pcsamp_warps_issue_stalled_barrier","smsp__pcsamp_warps_issue_stalled_branch_resolving","smsp__pcsamp_warps_issue_stalled_not_selected","smsp__pcsamp_warps_issue_stalled_math_pipe_throttle","smsp__pcsamp_warps_issue_stalled_mio_throttle","smsp__pcsamp_warps_issue_stalled_lg_throttle","smsp__pcsamp_warps_issue_stalled_no_instructions","smsp__pcsamp_warps_issue_stalled_dispatch_stall","smsp__pcsamp_warps_issue_stalled_selected","smsp__pcsamp_sample_buffer_full"]
for r in rows[2:]:
    print("-----")
    for i,h in enumerate(hdr):
        if h in want:
            print(f"{h:75s} {units[i]:12s} {r[i]}")
