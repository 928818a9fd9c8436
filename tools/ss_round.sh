#!/usr/bin/env bash
T=${1:-r02w}
O=gpurun_out/$T
mkdir -p $O
X=scenes/xml
timeout 600 python tools/ss_check.py $X/vader_wfpt_1080p.xml $X/vader_tex_wfpt_1080p.xml $X/dispersion_cbox_wfpt_1080p.xml scenes/ref_xml/whiskey.xml 2>&1 | tee $O/default.log
for lib in cuda-pt_b200/variants/*.so; do
  CPTB200_LIB=$PWD/$lib timeout 300 python tools/ss_check.py $X/vader_wfpt_1080p.xml 2>&1 | tee -a $O/variants.log
done
