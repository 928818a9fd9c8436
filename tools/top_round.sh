#!/usr/bin/env bash
T=${1:-r02top}
O=gpurun_out/$T
mkdir -p $O
V=$PWD/cuda-pt_b200/variants/libcpt_top1.so
CPTB200_LIB=$V timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $O/pytest_gpu_top1.log
for lib in cuda-pt_b200/libcpt_b200.so $V; do
  CPTB200_LIB=$lib timeout 200 python tools/auto_check.py scenes/xml/vader_wfpt_1080p.xml scenes/xml/vader_tex_wfpt_1080p.xml scenes/ref_xml/whiskey.xml 2>&1 | grep "^{" | tee -a $O/top_variants.log
done
