#!/usr/bin/env bash
# GPU-box session for the cooperative traversal: parity first, then the knob sweep for both group widths.
T=${1:-r02a}
O=gpurun_out/$T
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tee $O/gpu.txt
for W in 4 8; do
  CPTB200_COOP_WIDTH=$W timeout 600 python tools/coop_check.py parity 2>&1 | tee $O/parity_w$W.log
done
for W in 4 8; do
  CPTB200_COOP_WIDTH=$W timeout 900 python tools/coop_check.py bench 2>&1 | tee $O/bench_w$W.log
done
