#!/usr/bin/env bash
set -x
O=gpurun_out/${1:-ncuv2}
mkdir -p $O
# traversal over slot queues vs dense records, fog scene, same launch (bounce 5 of the first batch)
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 5 -c 1 -o $O/prof_fog_k_trace_slot python bench.py --scene fog --pipeline 1 --steps 1 --warmup 1 --no-cpu-baseline > $O/a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_vptd_trace -s 5 -c 1 -o $O/prof_fog_k_vptd_trace python bench.py --scene fog --steps 1 --warmup 1 --no-cpu-baseline > $O/b.log 2>&1
for k in k_vptd_post k_vptd_pre; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 5 -c 1 -o $O/prof_vpt_$k python bench.py --scene vpt --steps 1 --warmup 1 --no-cpu-baseline > $O/$k.log 2>&1
done
ls -la $O
