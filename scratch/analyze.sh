#!/bin/bash
# usage (here): scratch/analyze.sh <tag>  -- per-function profile + headline metrics of gpurun_out/prof_kstep_<tag>.ncu-rep
set -e
W=/tmp/cub_$1; rm -rf $W; mkdir -p $W; cd $W
cuobjdump -xelf all /root/repo/icebergs_b200/lib/libkid_b200.so > /dev/null
ncu -i /root/repo/gpurun_out/prof_kstep_$1.ncu-rep --page source --csv > src.csv 2>/dev/null
python /root/repo/profiles/line_profile.py src.csv kid_b200.sm_100a.cubin k_stepILb0ELb0ELb0ELb1ELb0E /root/repo/icebergs_b200/csrc | head -${2:-30}
ncu -i /root/repo/gpurun_out/prof_kstep_$1.ncu-rep --page raw --csv 2>/dev/null | python3 -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]; d=dict(zip(hdr,rows[2]))
for k in ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','sm__issue_active.avg.pct_of_peak_sustained_elapsed','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','sm__icc_request_hit_rate.pct','sass__inst_executed_local_loads','sass__inst_executed_local_stores','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio']:
    print(k, d.get(k))
"
