#!/usr/bin/env bash
# ncu --set full of the 5th launch (bounce 4) of the closest-hit and shadow kernels, per-thread (traversal 2) vs cooperative (4)
T=${1:-r02b}
O=gpurun_out/$T
mkdir -p $O
M="smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum"
for TR in 2 4; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 4 -c 1 -o $O/prof_trace_t$TR python tools/coop_check.py one $TR > $O/ncu_trace_t$TR.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shadow_apply -s 4 -c 1 -o $O/prof_shadow_t$TR python tools/coop_check.py one $TR > $O/ncu_shadow_t$TR.log 2>&1
done
ls -la $O
