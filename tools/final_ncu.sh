#!/usr/bin/env bash
# S sweep + ncu captures of the final traversal kernels (default per-thread kernels and the stream kernel)
T=${1:-r02v}
O=gpurun_out/$T
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tee $O/gpu.log
timeout 300 python tools/s_sweep.py 2>&1 | tee $O/s_sweep.log
CPTB200_SQ_REFILL=8 CPTB200_SQ_FLUSH=16 CPTB200_SQ_STEPS=4 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_traverse_stream --launch-skip 4 --launch-count 1 \
    -o $O/stream_final python tools/stream_check.py one 7 > $O/ncu_stream.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_shadow_apply" --launch-skip 8 --launch-count 2 \
    -o $O/base_final python tools/stream_check.py one 2 > $O/ncu_base.log 2>&1
tail -n 3 $O/ncu_stream.log $O/ncu_base.log
