#!/usr/bin/env bash
# Runs the UNMODIFIED reference (oracle/_ref/ref_dump) on the GPU box and writes golden scene dumps +
# float frames into gpurun_out/ref (copied from there into tests/golden/ref and committed).
set -x
mkdir -p gpurun_out/ref
R=oracle/_ref/ref_dump
X=scenes/xml
$R render $X/cbox_wfpt_256.xml gpurun_out/ref/cbox_wfpt_256 wfpt 0 1,2,4,16,64 > gpurun_out/ref/cbox_wfpt.log 2>&1
$R render $X/cbox_wfpt_256.xml gpurun_out/ref/cbox_wfpt_256_s7 wfpt 7 1,3 >> gpurun_out/ref/cbox_wfpt.log 2>&1
$R render $X/cbox_pt_256.xml gpurun_out/ref/cbox_pt_256 pt 0 1,64 > gpurun_out/ref/cbox_pt.log 2>&1
$R render $X/cbox_mats_wfpt_256.xml gpurun_out/ref/cbox_mats_256 wfpt 0 1,2,16,64 > gpurun_out/ref/cbox_mats.log 2>&1
$R render $X/vader_wfpt_256.xml gpurun_out/ref/vader_256 wfpt 0 1,8 > gpurun_out/ref/vader_256.log 2>&1
$R render $X/texroom_wfpt_256.xml gpurun_out/ref/texroom_256 wfpt 0 1,64 > gpurun_out/ref/texroom.log 2>&1
$R render $X/medium_cbox_vpt_256.xml gpurun_out/ref/medium_cbox_256 vpt 0 1,64 > gpurun_out/ref/medium.log 2>&1
ls -la gpurun_out/ref | head -60
$R render $X/grid_cbox_vpt_256.xml gpurun_out/ref/grid_cbox_256 vpt 0 1,16 > gpurun_out/ref/grid.log 2>&1
$R render $X/grid_cbox_rayleigh_vpt_256.xml gpurun_out/ref/grid_cbox_rayleigh_256 vpt 0 16 >> gpurun_out/ref/grid.log 2>&1
