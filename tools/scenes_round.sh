#!/usr/bin/env bash
T=${1:-r02sc}
O=gpurun_out/$T
mkdir -p $O
for sc in c2 whiskey c3; do
  timeout 120 python bench.py --scene $sc --steps 8 --warmup 4 --no-extras --no-cpu-baseline 2>&1 | tail -1 > $O/bench_${sc}_ours.json
  python -c "
import json; d=json.loads(open('$O/bench_${sc}_ours.json').read()); print('$sc', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms/step', d['config']['traversal'])"
done
