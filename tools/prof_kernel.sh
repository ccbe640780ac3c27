#!/bin/bash
# tools/prof_kernel.sh <kernel regex> <tag> [launch-skip] [extra bench args] -- one `ncu --set full` capture of one launch of one kernel inside a
# 32-session P step (after the same command ran to exit 0 without ncu), plus the per-source-line instruction table and the raw counters as text.
k=$1; tag=$2; skip=${3:-3}; shift 3 2>/dev/null
B="python bench.py --steps 3 --warmup 3 --sessions 32 --groups 1 --no-cpu --no-e2e $*"
mkdir -p gpurun_out
$B > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$tag.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip $skip -c 1 -f -o gpurun_out/prof_$tag $B > gpurun_out/ncu_$tag.log 2>&1
NCU_LINES_TOP=60 python tools/ncu_lines.py gpurun_out/prof_$tag.ncu-rep $k > gpurun_out/lines_$tag.txt 2> gpurun_out/lines_$tag.err
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/raw_$tag.csv 2>/dev/null
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/raw_$tag.csv")))
h,u,r=rows[0],rows[1],rows[2]
want=["gpu__time_duration.sum","smsp__inst_executed.sum","launch__registers_per_thread","sm__warps_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active","sm__throughput.avg.pct_of_peak_sustained_elapsed","smsp__issue_active.avg.pct_of_peak_sustained_active","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","dram__bytes_read.sum","dram__bytes_write.sum","smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct","smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct","smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct","smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct","smsp__warp_issue_stalled_not_selected_per_warp_active.pct","smsp__warp_issue_stalled_wait_per_warp_active.pct","smsp__warp_issue_stalled_barrier_per_warp_active.pct","smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct","launch__occupancy_limit_registers","launch__occupancy_limit_shared_mem","local_load_requests","smsp__inst_executed_op_local_ld.sum","smsp__inst_executed_op_local_st.sum"]
for w in want:
    if w in h:
        i=h.index(w); print(f"{w:90s} {r[i]} {u[i]}")
PY
head -45 gpurun_out/lines_$tag.txt
