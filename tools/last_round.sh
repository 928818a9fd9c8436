#!/usr/bin/env bash
O=gpurun_out/last; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/smoke.log
python bench.py --steps 12 --warmup 3 2>&1 | tail -1 > $O/bench_ours.json
python -c "
import json; j=json.load(open('$O/bench_ours.json')); print(j['value'], j['ms_per_step']/16, j['rays_per_sample'], j['e2e']['value'], j['gpu_launches'])"
