#!/usr/bin/env bash
# GPU-box session: reference frames of the Rayleigh / sggx scene, our renderer against them
set -x
O=gpurun_out/${1:-grid}
mkdir -p $O
R=oracle/_ref/ref_dump
timeout 300 $R render scenes/xml/grid_cbox_rayleigh_vpt_256.xml $O/grid_cbox_rayleigh_256 vpt 0 1,16 > $O/rayleigh.log 2>&1; tail -2 $O/rayleigh.log
timeout 300 python tools/grid_check.py $O grid_cbox_rayleigh_256 2>&1 | tee $O/check.log
