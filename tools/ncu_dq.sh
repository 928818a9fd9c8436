#!/usr/bin/env bash
T=${1:-r02l}
O=gpurun_out/$T
mkdir -p $O
CPTB200_DYN_REFILL=16 CPTB200_DYN_STEPS=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_traverse_dq -s 9 -c 1 -o $O/prof_trace_dq python tools/coop_check.py one 6 > $O/ncu_trace_dq.log 2>&1
tail -3 $O/ncu_trace_dq.log
timeout 300 python tools/coop_check.py one 2 2>&1 | tail -1 | tee $O/queue_counts.log
