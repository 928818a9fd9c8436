#!/usr/bin/env bash
T=${1:-r02sh}
O=gpurun_out/$T
mkdir -p $O
for lib in cuda-pt_b200/variants/*.so; do
  CPTB200_LIB=$PWD/$lib timeout 200 python tools/auto_check.py scenes/xml/vader_wfpt_1080p.xml 2>&1 | grep "^{" | tee -a $O/variants.log
done
