#!/usr/bin/env bash
set -x
O=gpurun_out/${1:-ncuv}
mkdir -p $O
for k in k_vpt_post k_vpt_pre k_vpt_march; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 5 -c 1 -o $O/prof_$k python bench.py --scene vpt --steps 1 --warmup 1 --no-cpu-baseline > $O/$k.log 2>&1
done
ls -la $O
