#!/bin/bash
# bottleneck counters of two k_trace launches (bounce 1 and 2) of one C3 frame (run under gpurun)
M=gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum,sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_local_op_st.sum,lts__t_sector_hit_rate.pct
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && ncu --metrics $M --clock-control none -k regex:"k_trace" -s 31 -c 2 --csv --log-file gpurun_out/trace_metrics.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu3.log 2>&1
python - <<'PY'
import csv
lines=[l for l in open("gpurun_out/trace_metrics.csv") if not l.startswith("==")]
d={}
for row in csv.DictReader(lines):
    d.setdefault(row["Metric Name"],[]).append(row["Metric Value"])
for k,v in d.items(): print(f"{k:95s}", v)
PY
