#!/usr/bin/env bash
# GPU round of the stream traversal: parity first (bounded), then the knob sweep with the default build and a short sweep per variant build
T=${1:-r02s}
O=gpurun_out/$T
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tee $O/gpu.log
timeout 300 python tools/stream_check.py parity 2>&1 | tail -20 | tee $O/parity.log
if ! grep -q "parity\] OK" $O/parity.log; then echo "PARITY FAILED - no sweep"; exit 1; fi
timeout 420 python tools/stream_check.py sweep 2>&1 | tee $O/sweep_default.log
for lib in cuda-pt_b200/variants/*.so; do
  echo "== $lib" | tee -a $O/sweep_variants.log
  CPTB200_LIB=$PWD/$lib timeout 200 python tools/stream_check.py bench 2>&1 | tee -a $O/sweep_variants.log
done
