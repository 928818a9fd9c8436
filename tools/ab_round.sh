#!/usr/bin/env bash
# A/B of the build variants under cuda-pt_b200/variants + the default build; optional ncu capture of the stream kernels
T=${1:-r02t}
O=gpurun_out/$T
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tee $O/gpu.log
timeout 300 python tools/ab_check.py 2>&1 | tee $O/ab_default.log
for lib in cuda-pt_b200/variants/*.so; do
  CPTB200_LIB=$PWD/$lib timeout 300 python tools/ab_check.py 2>&1 | tee -a $O/ab_variants.log
done
if [ "${NCU:-0}" = "1" ]; then
  CPTB200_SQ_REFILL=4 CPTB200_SQ_FLUSH=16 CPTB200_SQ_STEPS=8 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_traverse_stream --launch-skip 8 --launch-count 2 \
    -o $O/stream_r02t python tools/stream_check.py one 7 > $O/ncu_stream.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_shadow_apply" --launch-skip 8 --launch-count 2 \
    -o $O/base_r02t python tools/stream_check.py one 2 > $O/ncu_base.log 2>&1
  tail -3 $O/ncu_stream.log $O/ncu_base.log
fi
