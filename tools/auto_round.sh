#!/usr/bin/env bash
T=${1:-r02x}
O=gpurun_out/$T
mkdir -p $O
X=scenes/xml
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $O/pytest_gpu.log
timeout 600 python tools/auto_check.py $X/vader_wfpt_1080p.xml $X/vader_tex_wfpt_1080p.xml $X/dispersion_cbox_wfpt_1080p.xml scenes/ref_xml/whiskey.xml $X/cbox_wfpt_1080p.xml 2>&1 | tee $O/auto_default.log
timeout 300 python tools/knob_sweep.py 2>&1 | tee $O/knob_sweep.log
