#!/usr/bin/env bash
T=${1:-r02sr}
O=gpurun_out/$T
mkdir -p $O
for sc in c2 whiskey; do
  timeout 100 python bench.py --impl reference --scene $sc --steps 4 --warmup 1 --no-extras 2>&1 | tail -1 > $O/bench_${sc}_reference.json
  python -c "
import json; d=json.loads(open('$O/bench_${sc}_reference.json').read()); print('$sc', d.get('value'), 'Mrays/s', d.get('msamples_per_s'), 'Msamples/s')"
done
