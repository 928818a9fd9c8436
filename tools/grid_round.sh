#!/usr/bin/env bash
# GPU-box session for grid media: reference frames + scene dumps of the two grid scenes, then our renderer against them.
set -x
O=gpurun_out/${1:-grid}
mkdir -p $O
R=oracle/_ref/ref_dump
for s in grid_cbox grid_cbox_const; do
  timeout 300 $R render scenes/xml/${s}_vpt_256.xml $O/${s}_256 vpt 0 1,16,64 > $O/$s.log 2>&1; tail -2 $O/$s.log
done
ls -la $O
timeout 300 python tools/grid_check.py $O grid_cbox_256 grid_cbox_const_256 2>&1 | tee $O/check.log
