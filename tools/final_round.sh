#!/usr/bin/env bash
set -x
O=gpurun_out/${1:-final}
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee $O/smoke.log
python bench.py --steps 16 --warmup 3 2>&1 | tail -1 > $O/bench_ours.json
for w in vpt grid fog; do
  python bench.py --scene $w --steps 6 --warmup 3 2>&1 | tail -1 > $O/bench_${w}_ours.json
  python bench.py --scene $w --steps 6 --warmup 3 --pipeline 1 --no-cpu-baseline 2>&1 | tail -1 > $O/bench_${w}_ours_slot.json
done
python - <<PY
import json
for f in ("bench_ours","bench_vpt_ours","bench_vpt_ours_slot","bench_grid_ours","bench_grid_ours_slot","bench_fog_ours","bench_fog_ours_slot"):
    j=json.loads(open("$O/"+f+".json").read().strip().splitlines()[-1]); print(f, round(j["value"],1), round(j["ms_per_step"]/16,3), round(j["e2e"]["value"],1), j["roofline"]["kernel"], round(j["roofline"]["frac"],4), j["gpu_launches"])
PY
