#!/usr/bin/env bash
T=${1:-r02e}
O=gpurun_out/$T
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_coop -s 4 -c 1 -o $O/prof_trace_coop python tools/coop_check.py one 4 > $O/ncu_trace.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shadow_apply_coop -s 4 -c 1 -o $O/prof_shadow_coop python tools/coop_check.py one 4 > $O/ncu_shadow.log 2>&1
ls -la $O
